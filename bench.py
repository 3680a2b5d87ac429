#!/usr/bin/env python
"""bench.py -- Gcell-updates/s of the 2D monodomain time step (BASELINE.json's metric).

Workload (config.workload): configs[3] of BASELINE.json -- a 16384 x 16384 fibrillation
sheet (synthetic multi-spiral initial condition, reference default ionic model, standard PDE
mode, no-flux boundaries), Euler + 5-point Laplacian (the race-free, HBM-bound mode of
reactionDiffusion.cu), row-slab sharded over N GPUs with halo exchange ("scaling": "strong":
the sheet is fixed, N grows).  At N = 1 the whole sheet (8 GiB of FP64 state with its
ping-pong copy) lives on one B200 -- far larger than L2, so no flush is needed between steps.

One bench "step" = --substeps time steps of the whole sheet.

  python bench.py                                   # N=1
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference                  # the reference's own CUDA kernels (oracle/_ref)

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Gcell-updates/s"
BYTES_PER_UPDATE = 32.0   # read u,v + write u,v, FP64, once per time step


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", 1965.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()

    def run(self):
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.05)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4)
                          if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_params(yh, a):
    over = dict(timeIntOrder=1, lap4=0) if a.mode == "euler5" else {}
    return yh.default_params(a.nx, a.ny, scale_L=True, **over)


def timing_oracle(oracle_lib):
    """The plain-C oracle rebuilt for SPEED on this host (gcc -O3 -march=native -fopenmp;
    contraction allowed -- this build is only timed, parity uses the -ffp-contract=off one).  Built at
    run time because -march=native must match the box the baseline runs on; falls back to the parity
    build when no compiler is around."""
    src = os.path.join(ROOT, "oracle", "yh_oracle.c")
    out_dir = os.path.join(ROOT, "oracle", "_fast")
    so = os.path.join(out_dir, "libyh_oracle_fast.so")
    flags = "-O3 -march=native -fopenmp"
    try:
        os.makedirs(out_dir, exist_ok=True)
        subprocess.run(["gcc"] + flags.split() + ["-fPIC", "-std=c11", "-shared", "-o", so, src, "-lm"],
                       check=True, capture_output=True, timeout=120)
        return oracle_lib.Oracle(so), f"gcc {flags}, built on this host"
    except Exception:
        return oracle_lib.load(), "gcc -O2 -fopenmp -ffp-contract=off (parity build; no compiler at run time)"


def cpu_baseline(a, oracle_lib):
    """Plain-C oracle (OpenMP, all host cores) on a bounded sample of the same workload."""
    from yolohtli_b200 import synth
    o, how = timing_oracle(oracle_lib)
    n = min(a.nx, 4096)
    over = dict(timeIntOrder=1, lap4=0) if a.mode == "euler5" else {}
    p = o.params_default(n, n, scale_L=True, **over)
    u, v = synth.fibrillation_ic(n, n)
    o.rd_advance(p, 2, u, v)   # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    o.rd_advance(p, 4, u, v)   # calibration: size the sample for ~12 s of CPU work
    per_step = (time.perf_counter() - t0) / 4
    steps = int(max(8, min(4000, 12.0 / max(per_step, 1e-4))))
    t0 = time.perf_counter()
    o.rd_advance(p, steps, u, v)
    dt = time.perf_counter() - t0
    if dt < 8.0:   # the calibration steps ran cold and over-estimated the cost: resize from the real rate
        steps = int(max(steps, min(20000, steps * 12.0 / max(dt, 1e-3))))
        t0 = time.perf_counter()
        o.rd_advance(p, steps, u, v)
        dt = time.perf_counter() - t0
    return {"value": n * n * steps / dt / 1e9, "unit": METRIC, "cores": o.threads(), "kind": "port",
            "sample": f"{n}x{n} tile of the same fibrillation IC, {steps} steps, plain-C oracle "
                      f"({how}), wall clock {dt:.2f} s"}


def run_reference(a):
    """--impl reference: the UNMODIFIED reference kernels (reactionDiffusion_wrapper + swapSoA,
    main.cu:879-882) built headless for sm_100 with the reference's default flags, one GPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import oracle_lib
    from yolohtli_b200 import synth
    import yolohtli_b200 as yh
    p = make_params(yh, a)
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    base = {"metric": METRIC, "unit": METRIC, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference"}
    cfg = {"workload": f"{a.nx}x{a.ny} fibrillation sheet, {a.mode}, {a.substeps} time steps per bench step",
           "mode": a.mode, "nx": a.nx, "ny": a.ny, "substeps": a.substeps}
    if have_gpu and os.path.exists(oracle_lib.REF_SO):
        ref = oracle_lib.Reference(nofma=False)
        ref.init(p)
        u, v = synth.fibrillation_ic(a.nx, a.ny)
        sub = max(1, a.substeps // 4)   # bounded sample per step: the reference is several times slower
        for _ in range(max(a.warmup, 1)):
            ref.rd_run(u, v, 2, copy_back=False)
        ms = []
        for _ in range(a.steps):
            _, _, t = ref.rd_run(u, v, sub, copy_back=False)
            ms.append(t)
        tot = sum(ms) / 1e3
        val = a.nx * a.ny * sub * a.steps / tot / 1e9
        cfg["note"] = ("reference's own CUDA kernels (oracle/_ref/libyhref.so, nvcc -arch sm_100 -O3, default "
                       "fmad), kernel-only loop {reactionDiffusion_wrapper; swapSoA}, single GPU (the reference "
                       "has no multi-GPU path); timed with CUDA events inside the harness")
        out = dict(base, value=val, ms_per_step=tot / a.steps * 1e3, config=cfg,
                   cpu_baseline={"value": val, "unit": METRIC, "cores": 0, "kind": "reference",
                                 "sample": f"{sub} reference time steps per bench step on the GPU; state resident in HBM"},
                   e2e={"value": val, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   gpu_launches=0)
    else:
        cb = cpu_baseline(a, oracle_lib)
        cfg["note"] = "oracle/_ref or GPU not available: plain-C oracle port on the host cores"
        out = dict(base, value=cb["value"], ms_per_step=None, config=cfg, cpu_baseline=cb,
                   e2e={"value": cb["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   gpu_launches=0)
    print(json.dumps(out), flush=True)


def run_ours(a):
    import torch
    import torch.distributed as dist
    import yolohtli_b200 as yh
    from yolohtli_b200 import host, synth
    from yolohtli_b200.slab import SlabRunner

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- yolohtli_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    yh.load_library()
    p = make_params(yh, a)
    T = a.tb
    halo = T if a.mode == "euler5" else 4
    run = SlabRunner(p, rank=rank, world=world, halo=halo, device=dev, transport=a.transport)
    lay = run.lay
    u0, v0 = synth.fibrillation_ic(a.nx, a.ny, rows=(lay.g0, lay.g1))
    run.u[run.cur].copy_(torch.as_tensor(u0))
    run.v[run.cur].copy_(torch.as_tensor(v0))
    del u0, v0
    cells_total = a.nx * a.ny

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ----------------------------------------------------
    for _ in range(a.warmup):
        run.advance(a.substeps, tb=T)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        run.advance(a.substeps, tb=T)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if sampler else None
    value = cells_total * a.substeps * a.steps / (ms / 1e3) / 1e9
    # one launch per T time steps (Euler, temporally blocked) or per time step (fused RK4 + lap4 kernel)
    passes = -(-a.substeps // T) if a.mode == "euler5" else a.substeps
    launches = passes * a.steps

    # ---- end to end: pinned host -> device, substeps, device -> pinned host --------------
    own = lay.j1 - lay.j0
    hu = torch.empty((own, a.nx), dtype=torch.float64).pin_memory()
    hv = torch.empty((own, a.nx), dtype=torch.float64).pin_memory()
    uo, vo = run.owned()
    hu.copy_(uo)
    hv.copy_(vo)
    e2e_steps = max(1, min(a.steps, 2))

    def e2e_once(n=None):
        uo_, vo_ = run.owned()
        uo_.copy_(hu, non_blocking=True)
        vo_.copy_(hv, non_blocking=True)
        run.count = 0   # host data again: first pass treats it as raw
        run.advance(n or a.e2e_substeps, tb=T)
        uo2, vo2 = run.owned()
        hu.copy_(uo2, non_blocking=True)
        hv.copy_(vo2, non_blocking=True)

    e2e_once(a.substeps)
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_once()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_val = cells_total * a.e2e_substeps * e2e_steps / (float(ms2.item()) / 1e3) / 1e9
    checksum = float(hu.sum().item())

    if rank == 0:
        peak, peak_src, _ = peaks()
        # dominant kernel = the only kernel in the timed region; per-launch duration from the events above
        per_launch_ms = ms / launches
        cells_local = (lay.j1 - lay.j0) * a.nx
        steps_per_launch = T if a.mode == "euler5" else 1.0 / (passes / a.substeps)
        achieved = BYTES_PER_UPDATE * cells_local * steps_per_launch / (per_launch_ms / 1e3) / 1e9
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            traffic = json.load(open(tfile)).get(f"{a.mode}_tb{T}_{a.nx}x{a.ny}_n{world}")
        from tests import oracle_lib
        cb = cpu_baseline(a, oracle_lib) if world == 1 and not a.no_cpu_baseline else None
        out = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"BASELINE configs[3]: {a.nx}x{a.ny} fibrillation sheet, reference default ionic "
                            f"model, standard PDE mode, {a.mode}, row-slab sharded over {world} GPU(s); "
                            f"{a.substeps} time steps per bench step, {T} time steps per HBM pass",
                "mode": a.mode, "nx": a.nx, "ny": a.ny, "substeps": a.substeps, "tb_steps": T,
                "parallelism": f"slab{world}", "halo_rows": halo, "halo_transport": run.transport,
                "l2": "inputs (>= 1 GiB per array per GPU) are larger than the 126 MB L2; no flush needed",
                "arithmetic": "FP64, no FMA contraction (bit-identical to the plain-C oracle)",
                "checksum_u": checksum,
            },
            "e2e": {"value": e2e_val, "unit": METRIC, "h2d_bytes_per_step": 16 * cells_total,
                    "d2h_bytes_per_step": 16 * cells_total, "steps": e2e_steps,
                    "time_steps_per_call": a.e2e_substeps,
                    "note": "each e2e step = whole state pinned host -> HBM, e2e-substeps time steps (halo exchange "
                            "included), whole state HBM -> pinned host: the reference's use (upload once, main.cu:470; "
                            "run; download, main.cu:519)"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "rd_euler_stream" if a.mode == "euler5" else "rd_rk_stream",
                         "peak_source": peak_src,
                         "note": f"algorithmic 32 B per cell-update x {steps_per_launch:g} step(s) per launch; "
                                 f"temporal blocking lets frac exceed 1"},
            "clocks": clocks,
        }
        if cb:
            out["cpu_baseline"] = cb
        print(json.dumps(out), flush=True)
    run.close()
    if world > 1:
        dist.destroy_process_group()


def run_sweep(a):
    """BASELINE configs[4] (C5): independent 512 x 512 paced simulations, 32 per GPU, stimulation
    period swept 600 -> 100 ms over the 256 sheets of the full 8-GPU job, sAPD bookkeeping every
    step (contourMode == 1 loop, main.cu:879-885, 1035).  Replicas only: no data-path collective."""
    import torch
    import torch.distributed as dist
    import yolohtli_b200 as yh
    from yolohtli_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- yolohtli_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    yh.load_library()
    nx, nsim = 512, 32
    p = yh.default_params(nx, nx, timeIntOrder=1, lap4=0)
    periods = (np.linspace(600.0, 100.0, 256) / p.dt).astype(np.int32)
    mine = periods[(rank * nsim) % 256:(rank * nsim) % 256 + nsim]
    area = synth.stim_area_square(nx, nx)
    sim = yh.Sim(p, n_sims=nsim, device=local)
    pin = [torch.zeros((nsim, nx, nx), dtype=torch.float64).pin_memory().numpy() for _ in range(5)]
    zero = pin[0]                       # quiescent initial state; pinned host buffers for the e2e leg
    sim.set_state(zero, zero)
    sim.set_pacing(mine, int(10.0 / p.dt))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        sim.run_apd(a.substeps, stim_area=area)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        sim.run_apd(a.substeps, stim_area=None)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if sampler else None
    cells = nsim * nx * nx * world
    value = cells * a.substeps * a.steps / (ms / 1e3) / 1e9
    # end to end: host state in, paced steps + APD, host state and APD maps out
    e2e_steps = max(1, min(a.steps, 2))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sim.set_state(zero, zero)
        sim.run_apd(a.e2e_substeps, stim_area=None)
        sim.get_state(out=(pin[1], pin[2]))
        sim.get_apd(out=(pin[3], pin[4]))
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_val = cells * a.e2e_substeps * e2e_steps / float(dt.item()) / 1e9
    if rank == 0:
        peak, peak_src, _ = peaks()
        launches = -(-a.substeps // 4) * a.steps   # one fused {4 Euler steps + APD} launch per 4 time steps
        achieved = BYTES_PER_UPDATE * nsim * nx * nx * 4 / (ms / launches / 1e3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: batched sweep, {nsim * world} independent 512x512 paced "
                                   f"simulations ({nsim} per GPU, stimulation period 600 -> 100 ms), Euler + 5-point, "
                                   f"sAPD every step; {a.substeps} time steps per bench step; replicas only",
                       "mode": "euler5+sapd", "nx": nx, "ny": nx, "sheets_per_gpu": nsim, "substeps": a.substeps,
                       "parallelism": f"replicas{world}",
                       "l2": "32 sheets x 4 arrays x 2 MiB = 256 MiB of state per GPU, larger than the 126 MB L2",
                       "arithmetic": "FP64, no FMA contraction (bit-identical to the plain-C oracle)"},
            "e2e": {"value": e2e_val, "unit": METRIC, "h2d_bytes_per_step": 16 * nsim * nx * nx * world,
                    "d2h_bytes_per_step": 32 * nsim * nx * nx * world, "steps": e2e_steps,
                    "time_steps_per_call": a.e2e_substeps,
                    "note": "pinned host state -> device, e2e-substeps paced steps with APD bookkeeping, state + APD1/APD2 -> pinned host; wall clock"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "rd_euler_stream", "peak_source": peak_src,
                         "note": "algorithmic 32 B per cell-update x 4 steps per launch (APD state is touched only at threshold crossings)"},
            "clocks": clocks,
        }
        print(json.dumps(out), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=16384)
    ap.add_argument("--ny", type=int, default=16384)
    ap.add_argument("--mode", default="euler5", choices=["euler5", "rk4lap4"])
    ap.add_argument("--tb", type=int, default=4, help="time steps per HBM pass (1, 2, 4)")
    ap.add_argument("--substeps", type=int, default=64, help="time steps per bench step")
    ap.add_argument("--e2e-substeps", type=int, default=1024, help="time steps per host->device->host call")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="halo exchange: NVLink peer stores (CUDA IPC) or NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="sheet", choices=["sheet", "sweep"],
                    help="sheet: configs[3], one large sheet in row slabs (default, the headline); "
                         "sweep: configs[4], 32 independent 512^2 paced simulations per GPU")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "sweep":
        run_sweep(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
