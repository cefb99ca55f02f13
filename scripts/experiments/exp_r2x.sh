#!/bin/bash
# ncu --set full with source counters of the fused first-layer kernel (opt-in path)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
H3D_FUSE_C1=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_c1f" -s 6 -c 1 -o gpurun_out/r2x_c1f -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2x_ncu.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/r2x_c1f.ncu-rep
