#!/bin/bash
# 64 -> 64 channel specialisation: tests, then A/B of the bench (H3D_TC_C64=0 restores the generic kernel)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_tc_conv.py -q -m gpu --timeout 300 -x > gpurun_out/tc_c64.log 2>&1; echo "tc rc=$?"; tail -15 gpurun_out/tc_c64.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 600 -k "handsegnet or posenet_stage or lifting or full_pipeline" > gpurun_out/pipe_c64.log 2>&1; echo "pipe rc=$?"; tail -15 gpurun_out/pipe_c64.log
for cfg in "new H3D_X=1" "generic H3D_TC_C64=0"; do
  set -- $cfg; name=$1; shift
  for prec in bf16x3 fp16; do
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $prec > gpurun_out/exp_c64_${name}_$prec.json 2> gpurun_out/exp_c64_${name}_$prec.err; echo "bench $name $prec rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp_c64_${name}_$prec.json")); print("$name $prec", "img/s %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["by_class_ms_per_step"], d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
  done
done
