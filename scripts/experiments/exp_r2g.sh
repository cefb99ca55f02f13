#!/bin/bash
# round 2: elected-lane issue (uniform-register operands) in every role warp
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_tc_conv.py -q -m gpu --timeout 300 -x > gpurun_out/r2g_tc.log 2>&1; rc=$?; echo "tc rc=$rc"; tail -4 gpurun_out/r2g_tc.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_gpu_tc_conv.py > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2g_tests.log
for name in a b; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 2 --e2e-all-outputs 0 > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err; echo "$name rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2g_$name.json'));print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms sus', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'], d['roofline']['by_class_ms_per_step'])"
done
timeout 600 python bench.py --config 1 --steps 50 --warmup 5 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2g_c1.json 2> gpurun_out/r2g_c1.err
python -c "import json;d=json.load(open('gpurun_out/r2g_c1.json'));print('c1', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms')"
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1200 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:mask_grow" -c 1 -o gpurun_out/r2g_prof_mask -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2g_ncumask.log 2>&1; echo "ncumask rc=$?"
