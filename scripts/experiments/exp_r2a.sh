#!/bin/bash
# round 2: CTA-pair N-stacked kernel for Cout = 128 (H3D_TC_PAIR128) and 64-channel kernel on a CTA pair (H3D_TC_C64X2)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_tf_vectors.py -q -m gpu --timeout 300 > gpurun_out/r2a_tc.log 2>&1; echo "tc rc=$?"
tail -5 gpurun_out/r2a_tc.log
for cfg in "new:" "nopair128:H3D_TC_PAIR128=0" "noc64x2:H3D_TC_C64X2=0" "old:H3D_TC_PAIR128=0 H3D_TC_C64X2=0"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_$name.json 2> gpurun_out/r2a_$name.err; echo "$name rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2a_$name.json'));print('$name', round(d['value'],1), 'img/s', round(d['ms_per_step'],3),'ms', d['clocks']['sm_mhz'], d['roofline']['by_class_ms_per_step'])"
done
