#!/bin/bash
# mask grower v3 + graph guard: tests + bench + launch list
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py tests/test_gpu_properties.py tests/test_golden.py -q -m gpu --timeout 600 -x > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2m_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustain-seconds 2 --e2e-all-outputs 0 > gpurun_out/r2m_a.json 2> gpurun_out/r2m_a.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2m_a.json'));print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms sus', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1200 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
