#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2j_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
cat gpurun_out/r2j_bench.json | head -c 6000
for c in 1 2 3 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_c$c.json 2> gpurun_out/r2j_c$c.err; echo "config $c rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2j_c$c.json'));print('$c', round(d['value'],1), d['unit'], round(d['ms_per_step'],3),'ms', d['e2e'], d['clocks']['sm_mhz'], d['gpu_launches'])"
done
