#!/bin/bash
# fused conv1_1 + conv1_2 kernel, separate mid / final epilogue warps: parity subset, kernel timing, A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -x -k "first_layer or small_odd or fused_first" > gpurun_out/r2w_t1.log 2>&1; echo "t1 rc=$?"; tail -5 gpurun_out/r2w_t1.log
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
H3D_FUSE_C1=1 timeout 600 ncu --metrics $M --clock-control none -k "regex:conv_c1f" -c 8 --csv --log-file gpurun_out/r2w_c1f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv
rows=[l for l in open('gpurun_out/r2w_c1f.csv') if not l.startswith('==')]
r=list(csv.DictReader(rows))
print([(x['ID'], x['Metric Name'][:12], x['Metric Value']) for x in r if 'c1f' in x['Kernel Name']][-4:])
P
for v in 0 1 0 1; do
  H3D_FUSE_C1=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustain-seconds 2 --e2e-all-outputs 0 > gpurun_out/r2w_f$v.json 2> gpurun_out/r2w_f$v.err; echo "fuse $v rc=$?"; tail -2 gpurun_out/r2w_f$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2w_f$v.json'));print('fuse $v', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms sus', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"
done
