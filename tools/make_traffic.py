"""profiles/traffic.json and profiles/r2_kernel_summary.md from the exported raw pages of the ncu captures
(tools/gpu_r2_capture.sh): measured DRAM bytes per launch of the kernels bench.py times, and one table
with what decides each kernel (duration, registers, occupancy limiters, FP64 pipe, issue slots, shared-memory
data pipe, DRAM), so that every number quoted in DESIGN.md has a kept artefact."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2c"


def raw(name):
    path = os.path.join(PROF, f"{tag}_{name}_raw.csv")
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(d, key):
    v, u = d[key]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
             "Tbyte/s": 1e12, "Gbyte/s": 1e9}.get(u, 1.0)
    return x * scale


def mix(name):
    path = os.path.join(PROF, f"{tag}_{name}_mix.txt")
    txt = open(path).read()
    m = re.search(r"executed: ([\d,]+)\s+FP64: ([\d,]+).*?other: ([\d,]+)", txt)
    return [int(x.replace(",", "")) for x in m.groups()] if m else None


KERNELS = [  # capture, what it is, key for traffic.json (or None), cells x steps per launch
    ("euler_quad_tb4_16384", "rd_euler_quad<4,128> 16384^2, 4 steps per pass (bench headline)", "euler5_tb4_16384x16384_n1", 16384 ** 2 * 4),
    ("euler_quad_tb4_16384_fast", "same, FAST arithmetic", "euler5_tb4_16384x16384_n1_fast", 16384 ** 2 * 4),
    ("euler_quad_tb1_8192", "rd_euler_quad<1,128> 8192^2, TMA feed, 1 step per pass", "euler5_tb1_8192x8192_n1", 8192 ** 2),
    ("rk_quad_lap4_8192", "rd_rk_quad<lap4> 8192^2, one RK4 + lap4 step", "rk4lap4_tb1_8192x8192_n1", 8192 ** 2),
    ("rk_quad_lap4_8192_fast", "same, FAST arithmetic", "rk4lap4_tb1_8192x8192_n1_fast", 8192 ** 2),
    ("tile_rk_512", "rd_tile_rk<4,lap4> 512^2, one RK4 + lap4 step", None, 512 ** 2),
    ("tile_euler_512", "rd_tile_euler<4> 512^2, 4 steps", None, 512 ** 2 * 4),
]

traffic, src, lines = {}, {}, []
lines.append("| kernel | duration (ncu, cold) | regs | CTAs/SM limit (regs / smem) | warp instr. | FP64 | 2·FP64+other / issue slots | FP64 pipe | LSU shared pipe | DRAM bytes per launch | B per cell-update | DRAM % of ncu peak |")
lines.append("|---|---|---|---|---|---|---|---|---|---|---|---|")
for name, what, key, updates in KERNELS:
    try:
        d = raw(name)
    except Exception:
        continue
    dur = num(d, "gpu__time_duration.sum")
    rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
    cyc = num(d, "sm__cycles_elapsed.max")
    slots = cyc * 148 * 4
    mx = mix(name)
    model = f"{(2 * mx[1] + mx[2]) / slots:.2f}" if mx else "-"
    lines.append(f"| `{tag}_{name}`: {what} | {dur * 1e3:.3f} ms | {d['launch__registers_per_thread'][0]} | "
                 f"{float(d['launch__occupancy_limit_registers'][0]):.0f} / {float(d['launch__occupancy_limit_shared_mem'][0]):.0f} | "
                 f"{mx[0] / 1e6:.1f} M | {mx[1] / 1e6:.1f} M | {model} | "
                 f"{float(d['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'][0]):.1f} % | "
                 f"{float(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'][0]):.1f} % | "
                 f"{(rd + wr) / 1e9:.3f} GB | {(rd + wr) / updates:.2f} | "
                 f"{float(d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0]):.1f} % |")
    if key:
        traffic[key] = rd + wr
        src[key] = (f"profiles/{tag}_{name}_raw.csv (dram__bytes_read.sum {rd / 1e9:.3f} GB + dram__bytes_write.sum "
                    f"{wr / 1e9:.3f} GB per launch; {dur * 1e3:.3f} ms under ncu)")
traffic["_source"] = src
json.dump(traffic, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
open(os.path.join(PROF, f"{tag[:2]}_kernel_summary.md"), "w").write(
    f"# ncu summary of the timed kernels (capture set `{tag}`, tools/gpu_r2_capture.sh; B200, --clock-control none)\n\n"
    "Issue slots = sm__cycles_elapsed.max x 148 SMs x 4 schedulers; an FP64 instruction holds a scheduler's dispatch port\n"
    "two cycles (profiles/r2_issue_mix.txt), so `2*FP64 + other` is the issue time the executed instructions need.\n\n"
    + "\n".join(lines) + "\n")
print("\n".join(lines))
