#!/bin/bash
# fused conv1_1 + conv1_2 kernel, iteration 2: parity subset, A/B timing, launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -x -k "handsegnet or first_layer or full_pipeline" > gpurun_out/r2q_t1.log 2>&1; echo "t1 rc=$?"; tail -5 gpurun_out/r2q_t1.log
for v in 0 1 0 1; do
  H3D_FUSE_C1=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustain-seconds 2 --e2e-all-outputs 0 > gpurun_out/r2q_f$v.json 2> gpurun_out/r2q_f$v.err; echo "fuse $v rc=$?"; tail -2 gpurun_out/r2q_f$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2q_f$v.json'));print('fuse $v', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms sus', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"
done
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1200 ncu --metrics $M --clock-control none -k "regex:conv_(c1f|c64x2|c3_tma)" -c 40 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc=$?"
grep -c c1f gpurun_out/r2q_launches.csv
python - <<'P'
import csv
rows=[l for l in open('gpurun_out/r2q_launches.csv') if not l.startswith('==')]
r=list(csv.DictReader(rows))
for x in r[-40:]:
    if x['Metric Name'] in ('gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed') and 'c1f' in x['Kernel Name']:
        print(x['ID'], x['Kernel Name'][:30], x['Metric Name'][:30], x['Metric Value'])
P
