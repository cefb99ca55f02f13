#!/bin/bash
# final validation with the fused first layers as default: full suite, default bench, config 1, launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2y_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2y_bench.json'));print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), 'all', round(d['e2e_all_outputs']['value'],1), 'sus', round(d['sustained']['value'],1), d['clocks'], d['gpu_launches'], 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
timeout 600 python bench.py --config 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2y_c1.json 2> gpurun_out/r2y_c1.err; echo "config 1 rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2y_c1.json'));print('1', round(d['value'],1), d['unit'], round(d['ms_per_step'],3),'ms', d['gpu_launches'])"
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1200 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2y_ncu.log 2>&1; echo "ncu rc=$?"
