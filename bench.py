#!/usr/bin/env python
"""bench.py -- Gcell-updates/s of the 2D monodomain time step (BASELINE.json's metric).

Workload (config.workload): configs[3] of BASELINE.json -- a 16384 x 16384 fibrillation
sheet (synthetic multi-spiral initial condition, reference default ionic model, standard PDE
mode, no-flux boundaries), Euler + 5-point Laplacian (the race-free, HBM-bound mode of
reactionDiffusion.cu), row-slab sharded over N GPUs with halo exchange ("scaling": "strong":
the sheet is fixed, N grows).  At N = 1 the whole sheet (8 GiB of FP64 state with its
ping-pong copy) lives on one B200 -- far larger than L2, so no flush is needed between steps.

One bench "step" = --substeps x N time steps of the whole sheet (N = number of GPUs, so that the timed
region stays about a second at N = 8).  The step loop runs in the C++ slab driver of the library
(include/yolohtli_slab.h); Python only moves the 256-byte IPC handles at start-up.

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
        # NVML in-process (a sample every ~10 ms); nvidia-smi subprocesses (~0.2 s each) only as the fallback
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [(getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),),
                    (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),),
                    (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),),
                    (getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),)]
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop_ev.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = get_reasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append([str(sm), str(mx), f"{pw:.1f}"] + ["Active" if (r & b[0]) else "Not Active" for b in bits])
                self._stop_ev.wait(0.01)
            return
        except Exception:
            pass
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


WORKLOAD = ("BASELINE configs[3]: {nx}x{ny} fibrillation sheet, reference default ionic model, standard PDE mode, "
            "no-flux boundaries, {mode}")
# extra single-GPU measurements printed beside the headline ("modes"): name -> (nx, mode, tb, time steps)
MODES = {
    "euler5_tb4_16384_fast": (16384, "euler5", 4, 64),   # the headline workload in the FAST arithmetic flavour
    "euler5_tb1_8192": (8192, "euler5", 1, 32),      # one time step per HBM pass: the HBM-bound form of the step
    "rk4lap4_8192": (8192, "rk4lap4", 0, 12),        # the reference's DEFAULT mode (saveFiles.cu:124-132) on a large sheet
    "rk4lap4_8192_fast": (8192, "rk4lap4", 0, 12),   # ... in the FAST arithmetic flavour (yh_set_arithmetic, tests/test_gpu_arith.py)
    "rk4lap4_512": (512, "rk4lap4", 0, 2048),        # ... and on its default 512^2 sheet (BASELINE configs[0])
    "rk4lap4_512_fast": (512, "rk4lap4", 0, 2048),
    "euler5_512": (512, "euler5", 4, 8192),
    "euler5_holes_1024": (1024, "euler5holes", 4, 2000),   # BASELINE configs[1]: blood-vessel obstacle masks, fibrillation IC
    "rk4_holes_1024": (1024, "rk4holes", 0, 500),          # ... in the reference's default time integrator (no lap4 in the mask branch)
}
FP64_PER_UPDATE = {"euler5": (8, 24), "rk4lap4": (36, 316)}   # (DFMA, DADD + DMUL) warp-lane instructions per cell update


def workload_config(a):
    """The keys that define the workload: identical in the `ours` and `reference` arms."""
    return {"workload": WORKLOAD.format(nx=a.nx, ny=a.ny, mode=a.mode), "mode": a.mode, "nx": a.nx, "ny": a.ny}


def mode_params(params_default, nx, ny, mode):
    over = dict(timeIntOrder=1, lap4=0) if mode.startswith("euler5") else {}
    if mode.endswith("holes"):
        over["solidSwitch"] = 1
    return params_default(nx, ny, scale_L=True, **over)


def mode_inputs(synth, n, mode):
    """Initial condition (and obstacle mask) of a `modes` measurement: identical in both arms."""
    u, v = synth.fibrillation_ic(n, n) if n >= 1024 else synth.cross_field_ic(n, n)
    mask = None
    if mode.endswith("holes"):
        mask = synth.hole_mask(n, seed=1)
        u, v = u * mask, v * mask
    return u, v, mask


def fp64_peak():
    """FP64 instruction issue peaks of this GPU (tools/fp64_peak.cu, ~0.3 s): warp-lane instructions / s
    for DFMA and for separate DMUL / DADD."""
    exe = os.path.join(ROOT, "tools", "bin", "fp64_peak")
    try:
        if not os.path.exists(exe):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe,
                            os.path.join(ROOT, "tools", "fp64_peak.cu")], check=True, capture_output=True, timeout=120)
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        best = {}
        for line in out.splitlines():
            d = json.loads(line)
            for k, v in d.items():
                best[k] = max(best.get(k, 0.0), v["Ginstr_per_s"] * 1e9)
        return {"dfma_per_s": best["fp64_dfma"], "dmul_dadd_per_s": best["fp64_dmul_dadd"],
                "source": "measured now (tools/fp64_peak.cu)"}
    except Exception as e:   # noqa: BLE001
        return {"dfma_per_s": 16.99e12, "dmul_dadd_per_s": 18.6e12,
                "source": f"fallback: round-2 measurement on this pool (tools/fp64_peak.cu could not run: {e})"}


def fp64_fraction(mode, updates_per_s, pk):
    """Share of the FP64 issue time the arithmetic of the step needs (exact, contraction-free build)."""
    fma, other = FP64_PER_UPDATE[mode]
    return updates_per_s * (fma / pk["dfma_per_s"] + other / pk["dmul_dadd_per_s"])


# ---- BASELINE configs[2] (C3: symmetry-reduction step) and configs[4] (C5: batched restitution sweep) as `modes` ----
SR_STEPS, SWEEP_SHEETS, SWEEP_STEPS = 1000, 32, 600


def ours_sr_mode(yh, torch):
    """C3: 512^2 meandering spiral, every step = RD (RK4 + lap4) + tips + 12 phase-condition integrals + 3x3 solve +
    BFECC, device-resident loop (yh_sim_run_sr_device); wall clock around the call."""
    import time
    nx = 512
    sim = yh.Sim(yh.default_params(nx, nx, timeIntOrder=1, lap4=0))
    sim.cross_field_ic()
    sim.run(12001, tb_steps=4)          # the cross-field initial condition curls into a spiral
    tips = sim.tips()
    u0, v0 = sim.get_state()
    sim.close()
    tx, ty = (float(tips[-1]["x"]), float(tips[-1]["y"])) if len(tips) else (nx / 2.0, nx / 2.0)
    sim = yh.Sim(yh.default_params(nx, nx, reduce_sym=True, tipx0=tx, tipy0=ty))
    sim.set_state(u0, v0)
    sim.run_sr(2, record=False)
    sim.run_sr_device(64, record=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.run_sr_device(SR_STEPS, record=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sim.close()
    return {"value": nx * nx * SR_STEPS / dt / 1e9, "unit": METRIC, "ms_per_time_step": dt * 1e3 / SR_STEPS,
            "arithmetic": "exact", "note": "BASELINE configs[2]: one symmetry-reduction step per time step, no host round trip"}


def ours_sweep_mode(yh, torch, synth):
    """C5: 32 paced 512^2 sheets per GPU (of the 256 of configs[4]), Euler + 5-point, sAPD every step."""
    import time
    nx = 512
    p = yh.default_params(nx, nx, timeIntOrder=1, lap4=0)
    periods = (np.linspace(600.0, 100.0, 256)[:SWEEP_SHEETS] / p.dt).astype(np.int32)
    area = synth.stim_area_square(nx, nx)
    sim = yh.Sim(p, n_sims=SWEEP_SHEETS)
    sim.set_pacing(periods, int(10.0 / p.dt))
    sim.run_apd(20, stim_area=area)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim.run_apd(SWEEP_STEPS, stim_area=area)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sim.close()
    return {"value": SWEEP_SHEETS * nx * nx * SWEEP_STEPS / dt / 1e9, "unit": METRIC, "ms_per_time_step": dt * 1e3 / SWEEP_STEPS,
            "arithmetic": "exact", "note": f"BASELINE configs[4]: {SWEEP_SHEETS} sheets per GPU in one batch"}


def ref_sr_mode(ref, o, synth):
    """The reference's own symmetry-reduction loop (main.cu:894-954 through its wrappers) on a spiral grown with its
    own kernels; the disc centre is the last tip of its own tip kernel."""
    nx, nsr = 512, 300
    ref.init(o.params_default(nx, nx, timeIntOrder=1, lap4=0))
    u, v = synth.cross_field_ic(nx, nx)
    u1, v1, _ = ref.rd_run(u, v, 12000)
    u2, v2, _ = ref.rd_run(u1, v1, 1)
    tips = ref.tip(u2, u1)
    tx, ty = (float(tips[-1]["x"]), float(tips[-1]["y"])) if len(tips) else (nx / 2.0, nx / 2.0)
    ref.init(o.params_default(nx, nx, reduce_sym=True, tipx0=tx, tipy0=ty))
    ref.sr_run(u2, v2, 20)
    ms = ref.sr_run(u2, v2, nsr)[3]
    return {"value": nx * nx * nsr / ms / 1e6, "unit": METRIC, "ms_per_time_step": ms / nsr}


def ref_sweep_mode(ref, o, synth):
    """The reference runs one sheet at a time (its contourMode == 1 loop with sAPD every step)."""
    nx = 512
    p = o.params_default(nx, nx, timeIntOrder=1, lap4=0)
    ref.init(p)
    area = synth.stim_area_square(nx, nx)
    z = np.zeros((nx, nx))
    per, dur = int(600.0 / p.dt), int(10.0 / p.dt)
    ref.apd_run(z, z, 20, per, dur, area)
    ms = ref.apd_run(z, z, SWEEP_STEPS, per, dur, area)[4]
    return {"value": nx * nx * SWEEP_STEPS / ms / 1e6, "unit": METRIC, "ms_per_time_step": ms / SWEEP_STEPS,
            "note": "one sheet at a time"}


def run_reference(a):
    """--impl reference: the UNMODIFIED reference kernels (reactionDiffusion_wrapper + swapSoA,
    main.cu:879-882) built headless for sm_100 with the reference's default flags, one GPU.  Nothing of
    libyolohtli_b200.so is loaded in this arm: parameters come from the oracle's restatement of
    parameterSetup(), the initial condition from numpy."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import oracle_lib
    from yolohtli_b200 import synth   # numpy only
    o = oracle_lib.load()
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    base = {"metric": METRIC, "unit": METRIC, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference"}
    cfg = workload_config(a)
    if have_gpu and os.path.exists(oracle_lib.REF_SO):
        ref = oracle_lib.Reference(nofma=False)

        def ref_rate(n, mode, nsteps, reps):
            ref.init(mode_params(o.params_default, n, n, mode))
            u, v, mask = mode_inputs(synth, n, mode)
            ref.rd_run(u, v, 2, solid=mask, copy_back=False)
            ms = [ref.rd_run(u, v, nsteps, solid=mask, copy_back=False)[2] for _ in range(reps)]
            return n * n * nsteps * reps / (sum(ms) / 1e3) / 1e9, sum(ms) / reps

        ref.init(mode_params(o.params_default, a.nx, a.ny, a.mode))
        u, v = synth.fibrillation_ic(a.nx, a.ny)
        for _ in range(max(a.warmup, 1)):
            ref.rd_run(u, v, 2, copy_back=False)
        ms = [ref.rd_run(u, v, a.substeps, copy_back=False)[2] for _ in range(a.steps)]
        tot = sum(ms) / 1e3
        val = a.nx * a.ny * a.substeps * a.steps / tot / 1e9
        del u, v
        modes = {}
        if not a.no_modes:
            for name, (n, mode, _tb, nsteps) in MODES.items():
                if name.endswith("_fast"):
                    continue      # same reference number as the exact entry
                ns = max(4, nsteps // 4)
                r, ms1 = ref_rate(n, mode, ns, 2)
                modes[name] = {"value": r, "unit": METRIC, "ms_per_time_step": ms1 / ns}
            for name, fn in (("sr_step_512", ref_sr_mode), ("sweep_32x512", ref_sweep_mode)):
                try:
                    modes[name] = fn(ref, o, synth)
                except Exception as e:   # noqa: BLE001 -- an extra measurement must not cost the line
                    modes[name] = {"error": f"{type(e).__name__}: {e}"}
        out = dict(base, value=val, ms_per_step=tot / a.steps * 1e3, config=cfg,
                   impl_config={"substeps": a.substeps,
                                "note": "reference's own CUDA kernels (oracle/_ref/libyhref.so: the reference's .cu files compiled "
                                        "in place by oracle/build_ref.sh, nvcc -arch sm_100 -O3, default fmad), kernel-only loop "
                                        "{reactionDiffusion_wrapper; swapSoA}, single GPU at every --gpus N (the reference has no "
                                        "multi-GPU path and no CPU path); timed with CUDA events inside the harness"},
                   cpu_baseline={"value": val, "unit": METRIC, "cores": 0, "kind": "reference",
                                 "sample": f"{a.substeps} reference time steps per bench step on the GPU; state resident in HBM"},
                   e2e={"value": val, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   modes=modes, gpu_launches=0)
    else:
        cb = cpu_baseline(a, oracle_lib)
        out = dict(base, value=cb["value"], ms_per_step=None, config=cfg,
                   impl_config={"note": "oracle/_ref or GPU not available: plain-C oracle port on the host cores"},
                   cpu_baseline=cb,
                   e2e={"value": cb["value"], "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   gpu_launches=0)
    print(json.dumps(out), flush=True)


def measure_mode(yh, torch, n, mode, tb, nsteps):
    """One extra single-GPU measurement through the C ABI (yh_rd_advance), state resident in HBM."""
    from yolohtli_b200 import host, synth
    p = mode_params(yh.default_params, n, n, mode)
    u0, v0, mask = mode_inputs(synth, n, mode)
    solid = torch.as_tensor(mask).cuda() if mask is not None else None
    uA, vA = torch.as_tensor(u0).cuda(), torch.as_tensor(v0).cuda()
    uB, vB = torch.zeros_like(uA), torch.zeros_like(vA)
    ru, rv = host.rd_advance(p, max(4, (nsteps // 8) & ~3), uA, vA, uB, vB, tb_steps=tb, solid=solid)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ou, ov = (uB, vB) if ru is uA else (uA, vA)
        ru, rv = host.rd_advance(p, nsteps, ru, rv, ou, ov, tb_steps=tb, flags=host.RD_INPUT_CANONICAL, solid=solid)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None else min(best, t)
    return n * n * nsteps / (best / 1e3) / 1e9, best / nsteps


def run_ours(a):
    import torch
    import torch.distributed as dist
    import yolohtli_b200 as yh
    from yolohtli_b200 import synth
    from yolohtli_b200.slab import Slab

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
    p = mode_params(yh.default_params, a.nx, a.ny, a.mode)
    T = a.tb if a.mode == "euler5" else 1
    halo = T if a.mode == "euler5" else 4
    # a bench step = substeps time steps; scaled with N so that the timed region stays ~1 s at N = 8
    substeps = a.substeps * world

    # ---- the C++ slab driver (include/yolohtli_slab.h): one slab per rank, wired over CUDA IPC -------
    slab = Slab(p, rank, world, halo=halo, device=local)

    def gather_bytes(b):
        t = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [x.cpu().numpy().tobytes() for x in out]

    if world > 1:
        slab.connect_over(gather_bytes)
    u0, v0 = synth.fibrillation_ic(a.nx, a.ny, rows=(slab.g0, slab.g1))
    slab.set_state(u0, v0, with_ghosts=True)
    slab.sync()
    del u0, v0
    cells_total = a.nx * a.ny
    st = slab.stream()

    def barrier():
        slab.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def global_checksum():
        cu, cv = slab.checksum()
        parts = [(cu, cv)]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (cu, cv))
        return sum(q[0] for q in parts) % (1 << 64), sum(q[1] for q in parts) % (1 << 64)

    # ---- N-invariant record of the result: 64 time steps from the initial condition, then the checksum
    slab.advance(64, T)
    chk = global_checksum()

    # ---- device-resident throughput ----------------------------------------------------
    for _ in range(a.warmup):
        slab.advance(substeps, T)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for _ in range(a.steps):
        slab.advance(substeps, T)
    e1.record(st)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clocks = sampler.stop() if sampler else None
    value = cells_total * substeps * a.steps / (ms / 1e3) / 1e9
    # kernels of this rank in the timed region: one pass per T time steps (Euler) or per step (RK); with
    # neighbours a pass is {one or two bands on the edge stream, the exchange kernel, the interior}
    passes = -(-substeps // T) if a.mode == "euler5" else substeps
    per_pass = 1 if world == 1 else (2 + (1 if 0 < rank < world - 1 else 0) + 1)
    launches = passes * a.steps * per_pass

    # ---- end to end through the C ABI: pinned host -> device, time steps, device -> pinned host --------
    own = slab.j1 - slab.j0
    hu = torch.empty((own, a.nx), dtype=torch.float64).pin_memory()
    hv = torch.empty((own, a.nx), dtype=torch.float64).pin_memory()
    slab.get_state(out=(hu, hv))
    e2e_steps = max(1, min(a.steps, 2))
    slab.run_host(hu, hv, hu, hv, 4 * T, T)                 # warm-up of the copy path
    barrier()
    e0.record(st)
    for _ in range(e2e_steps):
        slab.run_host(hu, hv, hu, hv, a.e2e_substeps, T)    # yh_slab_run_host: H2D, steps (halo exchange included), D2H
    e1.record(st)
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_val = cells_total * a.e2e_substeps * e2e_steps / (float(ms2.item()) / 1e3) / 1e9
    slab.close()
    del hu, hv

    if rank == 0:
        peak, peak_src, _ = peaks()
        per_launch_ms = ms / (passes * a.steps)      # duration of one pass over the slab (its kernels overlap)
        cells_local = own * a.nx
        achieved = BYTES_PER_UPDATE * cells_local * T / (per_launch_ms / 1e3) / 1e9
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tfile):
            traffic = json.load(open(tfile)).get(f"{a.mode}_tb{T}_{a.nx}x{a.ny}_n{world}")
        pk64 = fp64_peak()
        kernel = ("rd_euler_quad" if T == 4 else "rd_euler_stream") if a.mode == "euler5" else "rd_rk_stream"
        roof = {"bound": "hbm" if (a.mode == "euler5" and T == 1) else "fp64", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "kernel": kernel, "peak_source": peak_src,
                "hbm_actual_frac": (traffic / (per_launch_ms / 1e3) / 1e9 / peak) if traffic else None,
                "fp64_frac": fp64_fraction(a.mode, value * 1e9 / world, pk64), "fp64_peak": pk64,
                "note": f"achieved = algorithmic 32 B per cell-update x {T} time step(s) per pass / pass duration (CUDA events), so "
                        f"temporal blocking lets frac exceed 1; hbm_actual_frac = measured DRAM bytes per pass (ncu, profiles/) / "
                        f"pass duration / peak; fp64_frac = share of the FP64 issue time the step's {sum(FP64_PER_UPDATE[a.mode])} "
                        f"FP64 instructions per update need at the measured issue peaks (the binding roof for T >= 2 and for RK4)"}
        modes = {}
        if world == 1 and not a.no_modes:
            for name, (n, mode, tb, nsteps) in MODES.items():
                fast = name.endswith("_fast")
                yh.lib().yh_set_arithmetic(1 if fast else 0)
                r, ms1 = measure_mode(yh, torch, n, mode, tb, nsteps)
                yh.lib().yh_set_arithmetic(0)
                modes[name] = {"value": r, "unit": METRIC, "ms_per_time_step": ms1,
                               "hbm_algorithmic_frac": r * BYTES_PER_UPDATE / peak,
                               "arithmetic": "fast (coefficients combined, FMA chains; rounding-level differences, "
                                             "tests/test_gpu_arith.py)" if fast else "exact"}
                if not fast and mode in FP64_PER_UPDATE:
                    modes[name]["fp64_frac"] = fp64_fraction(mode, r * 1e9, pk64)
            for name, fn in (("sr_step_512", lambda: ours_sr_mode(yh, torch)),
                             ("sweep_32x512", lambda: ours_sweep_mode(yh, torch, synth))):
                try:
                    modes[name] = fn()
                except Exception as e:   # noqa: BLE001 -- an extra measurement must not cost the line
                    modes[name] = {"error": f"{type(e).__name__}: {e}"}
        from tests import oracle_lib
        cb = cpu_baseline(a, oracle_lib) if world == 1 and not a.no_cpu_baseline else None
        out = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a),
            "impl_config": {
                "substeps": substeps, "tb_steps": T, "parallelism": f"slab{world}", "halo_rows": halo,
                "driver": "C++ row-slab driver (include/yolohtli_slab.h): one slab per rank, CUDA-IPC peer mappings, halo rows by "
                          "NVLink peer stores + release/acquire flags from one exchange kernel, CUDA-graph replay of block pairs",
                "l2": "inputs (>= 1 GiB per array per GPU) are larger than the 126 MB L2; no flush needed",
                "arithmetic": "FP64, no FMA contraction (bit-identical to the plain-C oracle)",
                "checksum": {"what": "sum of the uint64 views of all cells mod 2^64 over all ranks (order independent: equal "
                                     "for every N)", "time_steps": 64, "u": f"{chk[0]:016x}", "v": f"{chk[1]:016x}"},
            },
            "e2e": {"value": e2e_val, "unit": METRIC, "h2d_bytes_per_step": 16 * cells_total,
                    "d2h_bytes_per_step": 16 * cells_total, "steps": e2e_steps,
                    "time_steps_per_call": a.e2e_substeps,
                    "note": "each e2e step = yh_slab_run_host on every rank: owned rows pinned host -> HBM, ghost rows from the "
                            "neighbours, e2e-substeps time steps (halo exchange included), owned rows HBM -> pinned host: the "
                            "reference's use (upload once, main.cu:470; run; download, main.cu:519)"},
            "gpu_launches": launches,
            "roofline": roof,
            "modes": modes,
            "clocks": clocks,
        }
        if cb:
            out["cpu_baseline"] = cb
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the extra single-GPU mode measurements")
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
