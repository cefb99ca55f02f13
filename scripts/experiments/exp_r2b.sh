#!/bin/bash
# round 2: full GPU suite + mismatch report + all BASELINE configs + launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2b_tests.log
timeout 900 python scripts/mismatch_report.py > gpurun_out/r02_mismatch.json 2> gpurun_out/r02_mismatch.err; echo "mismatch rc=$?"
tail -3 gpurun_out/r02_mismatch.err
timeout 900 python bench.py > gpurun_out/r2b_bench_c4.json 2> gpurun_out/r2b_bench_c4.err; echo "bench c4 rc=$?"; tail -2 gpurun_out/r2b_bench_c4.err
for c in 1 2 3 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 1 > gpurun_out/r2b_bench_c$c.json 2> gpurun_out/r2b_bench_c$c.err; echo "bench c$c rc=$?"; tail -2 gpurun_out/r2b_bench_c$c.err
done
python - <<'PY'
import json
for c in (4,1,2,3,5):
    try:
        d=json.load(open('gpurun_out/r2b_bench_c%d.json'%c))
        print('c%d'%c, round(d['value'],1),'img/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],1), 'all', d['e2e_all_outputs'] and round(d['e2e_all_outputs']['value'],1), 'sus', d['sustained'] and round(d['sustained']['value'],1), 'launches', d['gpu_launches'], 'frac', d['roofline'] and round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
    except Exception as e: print('c%d'%c,'ERR',e)
PY
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 700 --csv --log-file gpurun_out/r2b_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
