#!/usr/bin/env bash
# Round-1 evidence: launch list + full capture of the dominant kernel for the default bench command,
# plus RK4+lap4 (reference default mode) numbers for ours and the reference.
set -x
mkdir -p gpurun_out
# (1) launch list of the default bench command (short)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 64 --csv --log-file gpurun_out/r1_launches.csv \
   python bench.py --steps 2 --warmup 3 --substeps 16 --e2e-substeps 16 --no-cpu-baseline > gpurun_out/r1_launches_bench.log 2>&1
# (2) full capture of the dominant kernel at the bench size
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rd_euler_stream -s 14 -c 1 -o gpurun_out/r1_rd_euler_stream_16384 -f \
   python bench.py --steps 1 --warmup 3 --substeps 16 --e2e-substeps 16 --no-cpu-baseline > gpurun_out/r1_full_bench.log 2>&1
# (3) default-mode (RK4 + lap4) numbers: ours vs reference, 8192^2
timeout 900 python bench.py --mode rk4lap4 --nx 8192 --ny 8192 --steps 3 --warmup 3 --substeps 8 --e2e-substeps 32 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_rk4_8192.json
cat gpurun_out/bench_rk4_8192.json
timeout 900 python bench.py --impl reference --mode rk4lap4 --nx 8192 --ny 8192 --steps 3 --warmup 1 --substeps 8 2>&1 | tail -1 > gpurun_out/bench_ref_rk4_8192.json
cat gpurun_out/bench_ref_rk4_8192.json
# (4) headline, full default command + reference arm
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.json; cat gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference 2>&1 | tail -1 > gpurun_out/bench_default_ref.json; cat gpurun_out/bench_default_ref.json
