#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_pipeline.py tests/test_gpu_configs.py -q -m gpu --timeout 600 -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2i_tests.log
for name in a b; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 2 --e2e-all-outputs 0 > gpurun_out/r2i_$name.json 2> gpurun_out/r2i_$name.err; echo "$name rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2i_$name.json'));print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms sus', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"
done
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1200 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
