#!/usr/bin/env bash
# Edge-band height of the slab driver at the per-GPU load of the 8-GPU run (2048 rows x 16384), on 2 GPUs.
B="--no-cpu-baseline --no-modes --steps 10 --warmup 3 --nx 16384"
run2() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$1 bench.py --gpus 2 $B --ny 4096 2>&1 | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', round(d['value'],1), 'Gcell/s  e2e', round(d['e2e']['value'],1), ' ms/step', round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; }
timeout 300 python bench.py $B --ny 2048 2>&1 | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1 2048 rows', round(d['value'],1), 'Gcell/s  ms/step', round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
run2 1 "N=2 band=default(128)"
YH_SLAB_BAND=256 run2 2 "N=2 band=256"
YH_SLAB_BAND=64 run2 3 "N=2 band=64"
YH_SLAB_BAND=32 run2 4 "N=2 band=32"
