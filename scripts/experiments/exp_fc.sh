#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_golden.py tests/test_gpu_tc_conv.py -q -m gpu --timeout 600 -x -k "lifting or pose_prior or golden or case0 or case1" > gpurun_out/pipe_fc.log 2>&1; echo "pipe rc=$?"; tail -4 gpurun_out/pipe_fc.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp_fc.json 2> gpurun_out/exp_fc.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/exp_fc.json")); print("img/s %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["roofline"]["by_class_ms_per_step"], d["gpu_launches"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv_(tc|c64)" -s 231 -c 17 --csv --log-file gpurun_out/fc_launch.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/fc_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/fc_launch.csv') if not l.startswith('==')]
for r in csv.DictReader(rows): print(r['ID'], r['Kernel Name'][15:50], r['Grid Size'], r['Metric Value'])
PY
