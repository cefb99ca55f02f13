#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_ops.py tests/test_gpu_pipeline.py -q -m gpu --timeout 600 -x > gpurun_out/multi.log 2>&1; echo "multi+ops+pipe rc=$?"; tail -25 gpurun_out/multi.log | cut -c1-300
for g in p2p nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --gather $g > gpurun_out/bench2_$g.json 2> gpurun_out/bench2_$g.err; echo "bench2 $g rc=$?"; tail -3 gpurun_out/bench2_$g.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench2_$g.json").read().strip().splitlines()[-1])
    print("$g", "img/s %.0f e2e %.0f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["config"]["collective"])
except Exception as e: print("$g failed", e)
PY
done
