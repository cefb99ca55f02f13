#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out; export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_properties.py -q -m gpu --timeout 600 -x > gpurun_out/pp.log 2>&1; echo "pipe+prop rc=$?"; tail -4 gpurun_out/pp.log
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/exp_$name.json").read().strip().splitlines()[-1])
    print("$name", "img/s %.0f e2e %.0f ms %.3f tc %.2f direct %.2f launches %d clk %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["by_class_ms_per_step"]["tc_conv"], d["roofline"]["by_class_ms_per_step"]["direct_conv"], d["gpu_launches"], d["clocks"]))
except Exception as e: print("$name failed", e); print(open("gpurun_out/exp_$name.err").read()[-1500:])
PY
}
run nograph --steps 20 --warmup 3
run graph --steps 20 --warmup 3 --cuda-graph 1
run b1 --steps 50 --warmup 5 --batch 1
run b1graph --steps 50 --warmup 5 --batch 1 --cuda-graph 1
run b8graph --steps 30 --warmup 5 --batch 8 --cuda-graph 1
run b64graph --steps 10 --warmup 3 --batch 64 --cuda-graph 1
run long --steps 200 --warmup 3 --cuda-graph 1
