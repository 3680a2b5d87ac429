#!/usr/bin/env bash
# One 8-GPU box: slab parity vs one GPU, the headline sheet at N = 8 with the two dependency schedules (A/B),
# a per-rank timeline of the block schedule, the rk4lap4 default mode on slabs, the C5 sweep at N = 8.
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_n8.sh TAG'
TAG=${1:-n8}
N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${N}.json
YH_SLAB_WAIT=exchange timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-substeps 64 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${N}_wait_exchange.json
YH_SLAB_TIMELINE=gpurun_out/${TAG}_timeline timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --e2e-substeps 64 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${N}_timeline_run.json
timeout 300 $TR --master-port 29516 bench.py --gpus $N --steps 5 --warmup 3 --workload sweep 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_sweep_n${N}.json
timeout 300 $TR --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --mode rk4lap4 --e2e-substeps 64 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_rk4lap4_n${N}.json
python - <<PY
import json
for f in ["bench_n$N", "bench_n${N}_wait_exchange", "bench_n${N}_timeline_run", "bench_n${N}_timeline_wait_exchange_run", "bench_sweep_n$N", "bench_rk4lap4_n$N"]:
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % f))
        print(f, round(d["value"], 1), d["unit"], "| e2e", round(d["e2e"]["value"], 1), "| ms/step", round(d["ms_per_step"], 2), "| clocks", d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
