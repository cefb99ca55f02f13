#!/bin/bash
# 1 -> N GPU weak-scaling sweep exactly as the driver launches it
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; export PYTHONPATH=$PWD
NMAX=${1:-8}
for n in 1 2 4 8; do
  [ $n -gt $NMAX ] && break
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  fi
  echo "n=$n rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open("gpurun_out/scale_%d.json"%n).read().strip().splitlines()[-1])
        base = base or d["value"]
        print("N=%d value %.0f img/s (eff %.3f) e2e %.0f ms %.3f launches %d clocks %s | %s" % (n, d["value"], d["value"]/(n*base), d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"], d["config"]["collective"][:60]))
    except Exception as e: print(n, "failed", e)
PY
