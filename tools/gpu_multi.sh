#!/usr/bin/env bash
# usage: gpu_multi3.sh N -- parity of the slab driver vs one GPU, then the sheet and sweep workloads
set -x
N=${1:-2}
mkdir -p gpurun_out
YH_TRANSPORT=p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_nccl_check.py 2>&1 | grep -E "slab check|Error|error|Traceback" | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n${N}.json
python -c "import json; d=json.load(open('gpurun_out/bench_n${N}.json')); print('N=$N sheet', round(d['value'],1), 'Gcell/s; e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --workload sweep 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_sweep_n${N}.json
python -c "import json; d=json.load(open('gpurun_out/bench_sweep_n${N}.json')); print('N=$N sweep', round(d['value'],1), 'Gcell/s; e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2))"
