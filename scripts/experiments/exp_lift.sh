#!/bin/bash
# lifting on tensor cores + side streams: tests, then A/B of the bench (H3D_LIFT_DIRECT / H3D_NO_SIDE_STREAM restore the old behaviour)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -q -m gpu --timeout 300 -k "stride2 or padded or identity or case0-bf16x3" > gpurun_out/tc_s2.log 2>&1; echo "tc_s2 rc=$?"; tail -15 gpurun_out/tc_s2.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_c_client.py -q -m gpu --timeout 600 -k "lifting or pose_prior or full_pipeline or c_client" > gpurun_out/pipe_lift.log 2>&1; echo "pipe_lift rc=$?"; tail -15 gpurun_out/pipe_lift.log
for cfg in "new" "direct H3D_LIFT_DIRECT=1" "nostream H3D_NO_SIDE_STREAM=1" "old H3D_LIFT_DIRECT=1 H3D_NO_SIDE_STREAM=1"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_lift_$name.json 2> gpurun_out/exp_lift_$name.err; echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp_lift_$name.json")); print("$name", "img/s %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["by_class_ms_per_step"], d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
done
