#!/bin/bash
# round 2: reader generators, first-layer TMA-store epilogue, c64x2 instruction-shape experiments
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r2c_tests.log
for cfg in "new:" "c3old:H3D_C3_TMA=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err; echo "$name rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2c_$name.json'));print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['clocks']['sm_mhz'], d['roofline']['by_class_ms_per_step'])"
done
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
for e in 0 1 2 3; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:conv_c64x2 --csv --log-file gpurun_out/r2c_c64x2_exp$e.csv python scripts/experiments/mb_conv.py 32 320 320 64 64 3 tc_exp=$e > gpurun_out/r2c_mb_$e.log 2>&1
done
timeout 300 ncu --metrics $M --clock-control none -k regex:conv_c3 -c 12 --csv --log-file gpurun_out/r2c_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2c_ncu_c3.log 2>&1
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r2c_c64x2_exp*.csv'))+['gpurun_out/r2c_c3.csv']:
    rows=[r for r in csv.DictReader(l for l in open(f) if not l.startswith('=='))]
    by={}
    for r in rows: by.setdefault(r['ID'],{})[r['Metric Name']]=r['Metric Value']
    last=list(by.values())[-2:]
    for v in last: print(f, {k.split('.')[0][-28:]:x for k,x in v.items()})
PY
