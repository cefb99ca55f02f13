#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_tc_conv.py tests/test_golden.py tests/test_gpu_pipeline.py -q -m gpu --timeout 600 -x -k "not drivers" > gpurun_out/pipe_stack.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/pipe_stack.log
for cfg in "stack H3D_X=1" "nostack H3D_TC_STACK=0" "stack2 H3D_X=1" "nostack2 H3D_TC_STACK=0"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_stack_${name}.json 2> gpurun_out/exp_stack_${name}.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp_stack_${name}.json")); print("$name", "img/s %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["by_class_ms_per_step"], d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
done
