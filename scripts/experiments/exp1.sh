#!/bin/bash
cd "$(dirname "$0")/../.."; mkdir -p gpurun_out; export PYTHONPATH=$PWD
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err; echo "$name rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/exp_$name.json"))
print("$name", "img/s %.0f ms %.2f tc %.2f direct %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["by_class_ms_per_step"]["tc_conv"], d["roofline"]["by_class_ms_per_step"]["direct_conv"]))
PY
}
run base A=1
run bn256 H3D_TC_BN=256
run chunk3 H3D_TC_CHUNK_KB=3
run chunk1 H3D_TC_CHUNK_KB=1
run nopool H3D_NO_POOL_FUSION=1

H3D_TC_CHUNK_KB=3 timeout 600 python scripts/debug_errors.py > gpurun_out/errors_chunk3.log 2>&1
grep -E "^(bf16x3|fp16x3)" gpurun_out/errors_chunk3.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 520 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "ncu rc=$?"
