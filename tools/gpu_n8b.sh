#!/usr/bin/env bash
# 8 GPUs: the driver's bench command (pipelined e2e), then the e2e leg alone with more pipeline levels and with the plain schedule
TAG=${1:-n8b}
N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${N}.json
YH_SLAB_PIPE_LEVELS=56 timeout 150 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${N}_levels56.json
YH_SLAB_PIPE=0 timeout 150 $TR --master-port 29514 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_n${N}_nopipe.json
python - <<PY
import json
for f in ["bench_n$N", "bench_n${N}_levels56", "bench_n${N}_nopipe"]:
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % f))
        print(f, round(d["value"], 1), "| e2e", round(d["e2e"]["value"], 1), "| clocks", d.get("clocks"), d["impl_config"]["checksum"]["u"])
    except Exception as e:
        print(f, "FAILED", e)
PY
