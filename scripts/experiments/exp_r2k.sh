#!/bin/bash
# layer chains: correctness (chain 0/1/2 bit-identical, parity suites) + A/B timing
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_properties.py tests/test_gpu_tc_conv.py tests/test_gpu_pipeline.py -q -m gpu --timeout 600 -x > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2k_tests.log
for c in 0 2 1 0 1; do
  H3D_TC_CHAIN=$c timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --sustain-seconds 2 --e2e-all-outputs 0 > gpurun_out/r2k_c$c.json 2> gpurun_out/r2k_c$c.err; echo "chain $c rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2k_c$c.json'));print('chain $c', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms sus', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"
done
for c in 0 1; do
  H3D_TC_CHAIN=$c timeout 600 python bench.py --config 1 --steps 50 --warmup 5 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2k_b1_c$c.json 2> gpurun_out/r2k_b1_c$c.err; echo "B1 chain $c rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2k_b1_c$c.json'));print('B1 chain $c', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms')"
done
