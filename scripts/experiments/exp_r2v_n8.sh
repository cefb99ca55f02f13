#!/bin/bash
# N GPUs (gpurun --gpus N): scaling bench only; the p2p / multimem gather is verified against NCCL inside bench.py
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
N=${NGPU:-8}
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --sustain-seconds 1 > gpurun_out/r2v_bench_$N.json 2> gpurun_out/r2v_bench_$N.err; echo "bench $N rc=$?"; tail -3 gpurun_out/r2v_bench_$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2v_bench_$N.json').read().splitlines()[-1])
print('N=$N', round(d['value'],1), 'img/s e2e', round(d['e2e']['value'],1), 'all', d['e2e_all_outputs'] and round(d['e2e_all_outputs']['value'],1), 'verified', d['gather_verified'], d['config']['collective'][:90], 'numa', d['config']['numa_bound_cpus'], d['clocks'])
P
