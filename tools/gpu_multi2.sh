#!/usr/bin/env bash
# usage: gpu_multi2.sh N -- p2p vs nccl transport: parity + bench
set -x
N=${1:-2}
mkdir -p gpurun_out
for tr in p2p nccl; do
YH_TRANSPORT=$tr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/slab_nccl_check.py 2>&1 | grep -E "slab check|Error|error|Traceback" | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --transport $tr 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n${N}_$tr.json
python -c "import json; d=json.load(open('gpurun_out/bench_n${N}_$tr.json')); print('N=$N $tr', round(d['value'],1), 'Gcell/s; e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],2))"
done
