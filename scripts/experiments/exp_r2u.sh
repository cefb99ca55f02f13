#!/bin/bash
# finalisation pass: full GPU suite, default bench (with CPU baseline), BASELINE configs, launch list, ncu --set full of the top kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2u_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2u_bench.json'));print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), 'all', round(d['e2e_all_outputs']['value'],1), 'sus', round(d['sustained']['value'],1), d['clocks'], d['gpu_launches'], 'frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
for c in 1 2 3 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_c$c.json 2> gpurun_out/r2u_c$c.err; echo "config $c rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2u_c$c.json'));print('$c', round(d['value'],1), d['unit'], round(d['ms_per_step'],3),'ms e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'], 'frac', round(d['roofline']['frac'],3))"
done
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1200 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2u_ncu.log 2>&1; echo "ncu rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:conv_(tc2|c64x2|c3_tma)" -s 141 -c 12 -o gpurun_out/r2u_prof_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2u_ncufull.log 2>&1; echo "ncufull rc=$?"; ls -la gpurun_out/r2u_prof_tc.ncu-rep
