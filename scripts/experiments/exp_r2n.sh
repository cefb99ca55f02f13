#!/bin/bash
# small-batch tile policy: bit-identity across batch cuts, B = 1 latency
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_gpu_properties.py tests/test_gpu_tc_conv.py tests/test_gpu_pipeline.py tests/test_golden.py -q -m gpu --timeout 600 -x > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2n_tests.log
for v in 0 1; do
  H3D_TC_SMALL_SPLIT=$v timeout 600 python bench.py --config 1 --steps 100 --warmup 10 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2n_b1_$v.json 2> gpurun_out/r2n_b1_$v.err; echo "B1 split $v rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2n_b1_$v.json'));print('B1 split $v', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['gpu_launches'])"
  H3D_TC_SMALL_SPLIT=$v timeout 600 python bench.py --batch 4 --steps 50 --warmup 10 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2n_b4_$v.json 2> gpurun_out/r2n_b4_$v.err; echo "B4 split $v rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2n_b4_$v.json'));print('B4 split $v', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['gpu_launches'])"
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustain-seconds 2 --e2e-all-outputs 0 > gpurun_out/r2n_a.json 2> gpurun_out/r2n_a.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/r2n_a.json'));print('bench', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms sus', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"
