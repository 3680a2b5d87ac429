#!/usr/bin/env bash
# usage: gpu_multi.sh N   -- NCCL slab parity + bench at N GPUs
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_nccl_check.py 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n$N.json
cat gpurun_out/bench_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N', round(d['value'],1), 'Gcell/s; e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2))"
