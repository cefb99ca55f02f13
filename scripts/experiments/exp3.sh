#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/all_gpu.log 2>&1; echo "all gpu tests rc=$?"; tail -6 gpurun_out/all_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2.json 2> gpurun_out/bench_2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench_2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads(open("gpurun_out/bench_%d.json"%n).read().strip().splitlines()[-1])
        print(n, "img/s %.0f e2e %.0f ms %.2f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
    except Exception as e: print(n, "failed", e)
PY
