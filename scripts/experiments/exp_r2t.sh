#!/bin/bash
# timing experiment (wrong results): fused first-layer kernel without the block-1 conversion in the epilogue warps
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
for e in 0 1; do
H3D_TC_EXP=$e timeout 600 ncu --metrics $M --clock-control none -k "regex:conv_c1f" -c 8 --csv --log-file gpurun_out/r2t_e$e.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --sustain-seconds 0 --e2e-all-outputs 0 > gpurun_out/r2t_ncu$e.log 2>&1; echo "ncu exp=$e rc=$?"
grep "c1f" gpurun_out/r2t_e$e.csv | tail -4 | cut -d, -f1,12-20 | cut -c1-200
done
