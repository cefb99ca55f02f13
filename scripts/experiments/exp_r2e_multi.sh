#!/bin/bash
# round 2, N GPUs (gpurun --gpus N): multi-GPU tests (p2p gather incl. ragged slots, two contexts in one process) + scaling bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
N=${NGPU:-2}
nvidia-smi -L > gpurun_out/r2e_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 600 > gpurun_out/r2e_multi.log 2>&1; echo "multi rc=$?"; tail -8 gpurun_out/r2e_multi.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --sustain-seconds 1 > gpurun_out/r2e_bench_1.json 2> gpurun_out/r2e_bench_1.err; echo "bench 1 rc=$?"
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 5 --sustain-seconds 1 > gpurun_out/r2e_bench_$n.json 2> gpurun_out/r2e_bench_$n.err; echo "bench $n rc=$?"; tail -3 gpurun_out/r2e_bench_$n.err
  fi
  python -c "import json;d=json.load(open('gpurun_out/r2e_bench_$n.json'));print('N=$n', round(d['value'],1), 'img/s e2e', round(d['e2e']['value'],1), 'all', d['e2e_all_outputs'] and round(d['e2e_all_outputs']['value'],1), 'verified', d['gather_verified'], d['config']['collective'][:60], 'numa', d['config']['numa_bound_cpus'])"
done
