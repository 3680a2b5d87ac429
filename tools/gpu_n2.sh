#!/usr/bin/env bash
# 2 real GPUs: C++ hosts across two devices (plain, pipelined run_host, symmetry reduction), IPC across processes, bench at N = 2
TAG=${1:-n2}
mkdir -p gpurun_out
for a in "2048 2048 2 403 euler 2" "2048 4096 2 403 euler 2 pipe" "512 1024 2 61 rk4lap4 2 pipe" "512 512 2 200 sr 2" "2048 2048 2 60 sr 2"; do
  timeout 100 yolohtli_b200/lib/yh_slab_driver $a 2>&1 | tail -2 | tee -a gpurun_out/${TAG}_drivers.txt
done
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/slab_ipc_check.py 2048 230 2>&1 | grep slab_ipc_check | tee -a gpurun_out/${TAG}_drivers.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n2.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n2.json')); print('N=2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'], d['impl_config']['checksum']['u'])" | tee -a gpurun_out/${TAG}_drivers.txt
