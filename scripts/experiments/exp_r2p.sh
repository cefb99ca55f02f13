#!/bin/bash
# fused conv1_1 + conv1_2 kernel: parity tests, then A/B timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -x -k "handsegnet or first_layer" > gpurun_out/r2p_t1.log 2>&1; echo "t1 rc=$?"; tail -25 gpurun_out/r2p_t1.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_properties.py tests/test_golden.py -q -m gpu --timeout 600 > gpurun_out/r2p_t2.log 2>&1; echo "t2 rc=$?"; tail -8 gpurun_out/r2p_t2.log
for v in 0 1 0 1; do
  H3D_FUSE_C1=$v timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustain-seconds 2 --e2e-all-outputs 0 > gpurun_out/r2p_f$v.json 2> gpurun_out/r2p_f$v.err; echo "fuse $v rc=$?"; tail -2 gpurun_out/r2p_f$v.err
  python -c "import json;d=json.load(open('gpurun_out/r2p_f$v.json'));print('fuse $v', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms sus', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"
done
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 1200 ncu --metrics $M --clock-control none -c 700 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
