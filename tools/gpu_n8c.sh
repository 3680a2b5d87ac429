#!/usr/bin/env bash
# 8 GPUs: band height of the slab schedule (YH_SLAB_BAND) A/B on the headline sheet
TAG=${1:-n8c}
N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29520
for band in 16 32 64; do
  port=$((port+1))
  YH_SLAB_BAND=$band timeout 120 $TR --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-substeps 64 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${N}_band${band}.json
  python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n${N}_band${band}.json')); print('band $band:', round(d['value'],1), d['clocks'], d['impl_config']['checksum']['u'])" | tee -a gpurun_out/${TAG}_band.txt
done
