#!/bin/bash
# round 2: FC-chain kernel, PDL, new HBM kernels, reader generators, first-layer TMA store, c64x2 experiments
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
# the risky new kernel first, under a short timeout
timeout 300 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 120 -x -k "lifting or pose_prior" > gpurun_out/r2d_lift.log 2>&1; rc=$?; echo "lifting rc=$rc"; tail -5 gpurun_out/r2d_lift.log
if [ $rc -ne 0 ]; then
  H3D_FC_CHAIN=0 timeout 300 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 120 -x -k "lifting or pose_prior" > gpurun_out/r2d_lift_nochain.log 2>&1; echo "lifting (no chain) rc=$?"; tail -5 gpurun_out/r2d_lift_nochain.log
  export H3D_FC_CHAIN=0
fi
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2d_tests.log
for cfg in "new:" "c3old:H3D_C3_TMA=0" "nopdl:H3D_PDL=0" "nochain:H3D_FC_CHAIN=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err; echo "$name rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2d_$name.json'));print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['clocks']['sm_mhz'], d['gpu_launches'], d['roofline']['by_class_ms_per_step'])"
done
for cfg in "c1:" "c1nopdl:H3D_PDL=0" "c1nochain:H3D_FC_CHAIN=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --config 1 --steps 50 --warmup 5 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err; echo "$name rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2d_$name.json'));print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['gpu_launches'])"
done
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
for e in 0 1 2 3; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:conv_c64x2 --csv --log-file gpurun_out/r2d_c64x2_exp$e.csv python scripts/experiments/mb_conv.py 32 320 320 64 64 3 tc_exp=$e > gpurun_out/r2d_mb_$e.log 2>&1
done
timeout 1200 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r2d_c64x2_exp*.csv')):
    rows=[r for r in csv.DictReader(l for l in open(f) if not l.startswith('=='))]
    by={}
    for r in rows: by.setdefault(r['ID'],{})[r['Metric Name']]=r['Metric Value']
    for v in list(by.values())[-1:]: print(f, {k.split('.')[0][-28:]:x for k,x in v.items()})
PY
