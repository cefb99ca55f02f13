#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
export H3D_TC_2CTA=1
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -q -m gpu --timeout 120 -x > gpurun_out/tc2.log 2>&1; echo "tc2 rc=$?"; tail -15 gpurun_out/tc2.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 600 -k "stage or full_pipeline" > gpurun_out/pipe2.log 2>&1; echo "pipe2 rc=$?"; tail -8 gpurun_out/pipe2.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp_$name.json"))
    print("$name", "img/s %.0f ms %.2f tc %.2f direct %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["by_class_ms_per_step"]["tc_conv"], d["roofline"]["by_class_ms_per_step"]["direct_conv"]))
except Exception as e: print("$name failed", e); print(open("gpurun_out/exp_$name.err").read()[-1500:])
PY
}
run one H3D_TC_2CTA=0
run two H3D_TC_2CTA=1
run two_fp16 H3D_TC_2CTA=1 H3D_PRECISION=fp16
run two_bn128 H3D_TC_2CTA=1 H3D_TC_BN=128
