#!/bin/bash
# round 2: L2 prefetch in the 64-channel pair kernel, faster mask grower; full tests; profiles
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2f_tests.log
for cfg in "new:" "nopf:H3D_TC_EXP=4" "new2:"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err; echo "$name rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2f_$name.json'));print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['clocks']['sm_mhz'], d['gpu_launches'], d['roofline']['by_class_ms_per_step'])"
done
timeout 900 python bench.py > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err; echo "bench c4 rc=$?"; tail -2 gpurun_out/r2f_bench_c4.err
for c in 1 2 3 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 1 > gpurun_out/r2f_bench_c$c.json 2> gpurun_out/r2f_bench_c$c.err; echo "bench c$c rc=$?"; tail -2 gpurun_out/r2f_bench_c$c.err
done
python - <<'PY'
import json
for c in (4,1,2,3,5):
    try:
        d=json.load(open('gpurun_out/r2f_bench_c%d.json'%c))
        print('c%d'%c, round(d['value'],1),'img/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],1), 'all', d['e2e_all_outputs'] and round(d['e2e_all_outputs']['value'],1), 'sus', d['sustained'] and round(d['sustained']['value'],1), 'launches', d['gpu_launches'], 'frac', d['roofline'] and round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
    except Exception as e: print('c%d'%c,'ERR',e)
PY
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1200 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:conv_(tc2|c64x2)" -s 135 -c 5 -o gpurun_out/r2f_prof_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2f_ncufull.log 2>&1; echo "ncufull rc=$?"; ls -la gpurun_out/r2f_prof_tc.ncu-rep
