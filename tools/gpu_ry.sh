#!/usr/bin/env bash
# chunk height of the streaming Euler kernel on a 2048-row slab (the per-GPU share of the 8-GPU run)
B="--no-cpu-baseline --no-modes --steps 10 --warmup 3 --nx 16384"
one() { timeout 300 python bench.py $B --ny $1 2>&1 | grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ny=$1 RY=${YH_FAST_RY:-auto}', round(d['value'],1), 'Gcell/s  ms/step', round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; }
one 2048; one 4096; one 8192
for ry in 1024 683 512 342 256 171 128 64; do YH_FAST_RY=$ry one 2048; done
