#!/bin/bash
# first layer on tensor cores + 64->128 patch-reuse: tests, then A/B of the bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_tc_conv.py tests/test_golden.py -q -m gpu --timeout 300 -x > gpurun_out/tc_c3.log 2>&1; echo "tc rc=$?"; tail -15 gpurun_out/tc_c3.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 600 -k "handsegnet or posenet_stage or lifting or full_pipeline or fp16_fast" > gpurun_out/pipe_c3.log 2>&1; echo "pipe rc=$?"; tail -15 gpurun_out/pipe_c3.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
for cfg in "new H3D_X=1" "c3ffma H3D_C3_FFMA=1"; do
  set -- $cfg; name=$1; shift
  for prec in bf16x3; do
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision $prec > gpurun_out/exp_c3_${name}_$prec.json 2> gpurun_out/exp_c3_${name}_$prec.err; echo "bench $name $prec rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp_c3_${name}_$prec.json")); print("$name $prec", "img/s %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["by_class_ms_per_step"], d["gpu_launches"])
except Exception as e: print("$name failed", e)
PY
  done
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k "regex:conv_(c3|c64|tc_kernel)" -s 0 -c 12 --csv --log-file gpurun_out/c3_launch.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c3_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/c3_launch.csv') if not l.startswith('==')]
by={}
for r in csv.DictReader(rows):
    by.setdefault(r['ID'],{'n':r['Kernel Name'][:40]})[r['Metric Name']]=r['Metric Value']
for k,v in by.items(): print(k, v)
PY
