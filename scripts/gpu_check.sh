#!/bin/bash
# Runs on the GPU box (gpurun): every stage in its own process under a timeout, logs into gpurun_out/.
# usage: scripts/gpu_check.sh [stage ...]   stages: ops tc pipe bench ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
STAGES="${@:-ops tc pipe bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for s in $STAGES; do
  case $s in
    ops)   timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 300 > gpurun_out/ops.log 2>&1; echo "ops rc=$?" ;;
    tc)    timeout 900 python -m pytest tests/test_gpu_tc_conv.py -q -m gpu --timeout 300 > gpurun_out/tc.log 2>&1; echo "tc rc=$?" ;;
    pipe)  timeout 1500 python -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 600 > gpurun_out/pipe.log 2>&1; echo "pipe rc=$?" ;;
    smoke) timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench) for p in ${BENCH_PRECS:-bf16x3 fp16 fp32_ffma}; do
             timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --precision $p --no-cpu-baseline > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err; echo "bench $p rc=$?"
           done ;;
    benchfull) timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "benchfull rc=$?" ;;
    errors) timeout 900 python scripts/debug_errors.py > gpurun_out/errors.log 2>&1; echo "errors rc=$?" ;;
    sanitize) for tool in ${SAN_TOOLS:-memcheck synccheck}; do
             timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc_conv.py tests/test_golden.py tests/test_gpu_pipeline.py -q -m gpu -x --timeout 1200 \
               -k "(not test_conv2d_tc_vs_oracle and not pipeline and not stage and not network and not drivers and not fp16_fast and not tuple) or (test_conv2d_tc_vs_oracle and (case0 or case1 or case3 or case5 or case7 or case9)) or lifting_stage" > gpurun_out/sanitize_$tool.log 2>&1; echo "sanitize $tool rc=$?"; tail -5 gpurun_out/sanitize_$tool.log
           done ;;
    ref)   timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" ;;
    ncu)   timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
             python bench.py --steps 1 --warmup 3 --batch ${NCU_BATCH:-32} --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" ;;
    ncudram) timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k "regex:conv_(tc|c64)" -s ${TC_SKIP:-186} -c ${TC_COUNT:-62} --csv --log-file gpurun_out/tc_dram.csv \
             python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncudram.log 2>&1; echo "ncudram rc=$?" ;;
    ncufull) timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:conv_(tc|c64)" -s ${NCU_SKIP:-186} -c ${NCU_COUNT:-11} -o gpurun_out/prof_tc -f \
             python bench.py --steps 1 --warmup 3 --batch ${NCU_BATCH:-32} --no-cpu-baseline > gpurun_out/ncufull.log 2>&1; echo "ncufull rc=$?" ;;
  esac
done
tail -n 25 gpurun_out/*.log 2>/dev/null | tail -n 120
cat gpurun_out/bench_*.json 2>/dev/null
