#!/usr/bin/env python
"""Benchmark of the Veritas 1D1P Vlasov advance on B200 (metric of BASELINE.json):
phase-space cell-updates/s per RK stage = cells (both species) x 6 stages x steps / time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5] [--scaling strong|weak]

N = 1 : config 3 of BASELINE.json (uniform 65536 x 4096, single level, two species) — the largest single-GPU config.
N > 1 : config 5 (uniform 262144 x 4096 split in x over the GPUs, strong scaling); launched by torchrun, one rank per GPU.
--workload c1 | c2 | c4 : BASELINE.json's CPU-runnable configurations (2048 x 256 single level; 2 levels with the refinement in the
high-momentum tail; 3 levels with a regrid every 22 steps) driven through the C++ host classes (oracle/_ref/host_harness: the
reference's own driver loop compiled against veritas_b200/host), with the reference timed on the SAME configuration beside it.
One "step" = SolverManager::Advance(dt): 6 RK stages of moments + Poisson + fused Vlasov stage (x2 species) + Maxwell.
Inputs (f, fields) are resident in HBM for `value`; `e2e` drives the same steps through the host-facing API with the
per-step host round trips of the reference's driver loop (CalculateDt -> 8 B D2H, laser boundary values + dt -> H2D,
and the 1-D diagnostic arrays of fileOutput -> D2H) inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, np) ; two species with equal np
    "c1": (2048, 256),
    "c3": (65536, 4096),
    "c5": (262144, 4096),
}
# CPU arms (reference arm, cpu_baseline): a bounded sample of config 3 the reference can hold — config 3's p grid (4096 cells per
# species) on 2048 of its 65536 columns (its dense N x N Poisson matrix is 34 GB at N = 65536 and its patch storage 97 GB per
# species; at N = 2048: 32 MB and 6 GB), same case file, same per-cell work; ~2 s per step on 16 cores
REF_SAMPLE = (2048, 4096)
REF_SAMPLE_TEXT = ("sample of config 3: 2048 of its 65536 columns x 4096 p-cells, 2 species, single level (the reference's dense "
                   "N x N Poisson matrix and 360 B/cell storage do not fit configs 3/5)")
DENSITY = 0.1   # "laser pulse in underdense plasma" (BASELINE.json configs[0]); one Settings number (veritas.cpp:47)
# BASELINE.json configs 1, 2, 4 as harness arguments (nx np Lfinest + keys): the C++ host classes run them
HOST_WORKLOADS = {
    "c1": (["2048", "256", "1"], [], "c1: uniform 2048x256 x-p mesh, single level, 2 species"),
    "c2": (["1024", "128", "2"], ["refine_mode=1", "tail_p0=2", "regrid_every=22"],
           "c2: 1024x128 coarse mesh, 2 levels, refinement forced into the high-momentum tail, regrid every 22 steps, 2 species"),
    "c4": (["512", "64", "3"], ["regrid_every=22"], "c4: 512x64 coarse mesh, 3 levels, regrid every 22 steps (veritas.cpp:146-151), 2 species"),
}


# version of k_fused_stage the committed ncu traffic capture (profiles/fused_traffic.json) must match to be quoted as `traffic`
FUSED_KERNEL_VERSION = "v18"


def stage_bytes(cells_species, s):
    """Algorithmic bytes of one fused-stage launch: reads f^n (8 B; at s = 0 the same array as
    f^(s)), f^(s) (8), 2 s stored fluxes; writes f^(s+1) (8) and, except at s = 5, the new flux pair (16)."""
    return cells_species * (8 + (8 if s > 0 else 0) + 16 * s + 8 + (16 if s < 5 else 0))


def stage_bytes_moved(cells_species, s):
    """Bytes the shipped kernel moves by design: the algorithmic bytes plus the stage-0 low-order flux pair (written at s = 0,
    read at s >= 1: +16 B per cell) minus the history pair stage 5 skips (tableau weight b_1 = 0: -16 B per cell)."""
    return stage_bytes(cells_species, s) + cells_species * (16 - (16 if s == 5 else 0))


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """index of the next sample: brackets a region of interest inside a longer sampling run"""
        return len(self.rows)

    def stop(self, lo=None, hi=None):
        """lo, hi = mark() at the start / end of the timed region.  The sampler is started before the warm-up (nvidia-smi needs up
        to a second to deliver its first sample, longer than a 10-step timed region on 8 GPUs) and stopped after the end-to-end
        loop; samples inside [lo, hi] are used if there are any, otherwise all samples of the run — warm-up, timed and end-to-end
        steps, every one of them under the same load — and `window` says which."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        rows, window = self.rows, "whole run"
        if lo is not None and hi is not None:
            if hi > lo:
                rows, window = self.rows[lo:hi], "timed region"
            else:
                rows, window = self.rows, "warm-up + timed + end-to-end steps (the timed region is shorter than nvidia-smi's first-sample latency)"
        sm, mx, reasons = [], None, set()
        for r in rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}



def run_harness(exe, wl, steps, warmup, threads=None):
    """One timed run of the reference's driver loop (CalculateDt + Advance per step, regrids outside the timed region) in the
    harness build `exe`; returns the ORACLE_TIMING fields."""
    size, keys, _ = HOST_WORKLOADS[wl]
    cores = threads or os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([exe, "/dev/null"] + size + [str(DENSITY), str(steps)] + keys + ["time_only=1", f"warmup={warmup}"],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("ORACLE_TIMING")][-1]
    kv = dict(x.split("=") for x in line.split()[1:])
    return {"value": float(kv["cell_updates_per_s_per_stage"]), "seconds": float(kv["advance_s"]), "steps": int(kv["steps"]),
            "cells": int(kv["cells"]), "launches": int(kv.get("gpu_launches", 0)), "cores": cores}


def host_workload_line(args, impl):
    """BASELINE.json configs 1, 2, 4 through the reference's own class API: impl "ours" = oracle/_ref/host_harness (the case
    file compiled against veritas_b200/host + libveritas_b200.so), impl "reference" = oracle/_ref/ref_harness (against the
    unmodified reference, all host cores).  Both time CalculateDt + Advance of every step from t = 3T; f stays resident on the
    device between steps exactly as it stays in Rectangle::f, so the per-step host traffic of the GPU build is the 8-byte CFL
    bound down and dt + 12 laser values up: value and e2e are the same measurement."""
    wl = args.workload
    text = HOST_WORKLOADS[wl][2] + f", laser-plasma case n={DENSITY} N_c, t>=3T"
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    host = os.path.join(ROOT, "oracle", "_ref", "host_harness")
    cb = None
    if os.path.exists(ref) and not args.no_cpu_baseline:
        ref_steps = args.steps if impl == "reference" else min(args.steps, 22)
        r = run_harness(ref, wl, ref_steps, 1 if impl == "reference" else 0)
        cb = {"value": r["value"], "unit": "cell-updates/s", "cores": r["cores"], "kind": "reference", "same_config": True,
              "sample": f"{text}; {r['steps']} steps of CalculateDt + SolverManager::Advance, OMP_NUM_THREADS={r['cores']}"}
    if impl == "reference":
        if cb is None:
            return {"impl": "reference", "unavailable": "oracle/_ref/ref_harness missing (run __graft_entry__.build() in the build container)"}
        return {"impl": "reference", "metric": "phase-space cell-updates/s per RK stage", "value": cb["value"], "unit": "cell-updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": text}, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    if not os.path.exists(host):
        raise SystemExit("bench.py: oracle/_ref/host_harness missing (run __graft_entry__.build())")
    sampler = ClockSampler(0)
    sampler.start()
    g = run_harness(host, wl, args.steps, max(args.warmup, 3), threads=4)
    clocks = sampler.stop()
    return {"metric": "phase-space cell-updates/s per RK stage", "value": g["value"], "unit": "cell-updates/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * g["seconds"] / g["steps"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": text + f"; {g['cells']} patch cells at the end; small meshes: an L2 flush is not applied "
                                          "(the whole hierarchy is L2-resident in the reference-sized case; stated, not hidden)",
                       "parallelism": "single GPU", "api": "C++ host classes (SolverManager::CalculateDt / Advance / reGrid)", "cuda_graph": True},
            "clocks": clocks,
            "e2e": {"value": g["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 104, "d2h_bytes_per_step": 8,
                    "what": "the same measurement: the host classes' driver loop, host wall clock around CalculateDt + Advance"},
            "gpu_launches": g["launches"],
            "roofline": None, "roofline_note": "launch-latency-bound hierarchy of small patches (split path); the roofline figures are quoted on config 3",
            "cpu_baseline": cb}


def reference_arm(args, rank):
    """The reference's own CPU implementation (oracle/_ref/ref_harness = unmodified reference sources) on the host
    cores.  It cannot run configs 3/5 (dense N x N Poisson matrix: 34 GB / 550 GB, EMSolver.cpp:30-31), so each step is
    a step of REF_SAMPLE — config 3's p grid on 2048 of its columns, the same case file.  At most 40 steps are run (the value
    is a rate)."""
    if rank != 0:
        return
    if args.workload in HOST_WORKLOADS:
        line = host_workload_line(args, "reference")
        print(json.dumps(line))
        return
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    nx, np_ = REF_SAMPLE
    steps = min(40, max(1, args.steps + args.warmup))
    cores = os.cpu_count() or 1
    if not os.path.exists(harness):
        # the compiled reference did not travel: fall back to the C restatement (kind "port") is not implemented here
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness missing (run __graft_entry__.build() in the build container)"}))
        return
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([harness, "/dev/null", str(nx), str(np_), "1", str(DENSITY), str(steps), "time_only=1"],
                         env=env, stdout=subprocess.PIPE, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("ORACLE_TIMING")][-1]
    kv = dict(x.split("=") for x in line.split()[1:])
    value = float(kv["cell_updates_per_s_per_stage"])
    sec = float(kv["advance_s"])
    text = REF_SAMPLE_TEXT if args.gpus == 1 else REF_SAMPLE_TEXT.replace("sample of config 3: 2048 of its 65536", "sample of config 5: 2048 of its 262144")
    sample = f"{text}; {steps} steps of SolverManager::Advance from t=3T, OMP_NUM_THREADS={cores}"
    print(json.dumps({
        "impl": "reference", "metric": "phase-space cell-updates/s per RK stage", "value": value, "unit": "cell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / steps,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": sample},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def measure_fp64_peak():
    """fp64 issue peak (warp-wide thread-instructions per second) of this device: tools/fp64_peak run now, else the committed
    round-1 measurement"""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, check=True, timeout=60).stdout
        return float(json.loads(out.strip().splitlines()[-1])["fp64_instr_per_s"]), "tools/fp64_peak (DFMA micro-benchmark) run in this session"
    except Exception:
        return float(json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r1.json")))["fp64_instr_per_s"]), "profiles/fp64_peak_r1.json (round 1)"


def fused_vs_split_self_check(vb, S, np_, device):
    """Pre-flight check of the kernels this run is about to time, on this device: 256 columns of the workload's own p grid (the
    interior specialisation of the stage kernel and the long-column moments kernel), two free-running steps of the fused streaming
    path against the bit-faithful split path (which the tests pin to the reference).  Aborts the bench if they disagree."""
    import numpy as np
    res = {}
    for path in (S.PATH_SPLIT, S.PATH_FUSED):
        run = vb.LaserPlasmaRun(256, np_, density=DENSITY, device=device, path=path)
        run.init_device()
        run.time = 3 * run.T
        for _ in range(2):
            run.advance(run.calculate_dt())
        res[path] = [run.ctx.download_f(s, 0, 1) for s in range(2)] + [run.ctx.get_1d(S.J)]
        run.ctx.close()
    errs = [float(np.linalg.norm((b - a).ravel()) / max(np.linalg.norm(a.ravel()), 1e-300)) for a, b in zip(res[S.PATH_SPLIT], res[S.PATH_FUSED])]
    if not (errs[0] < 1e-12 and errs[1] < 1e-12 and errs[2] < 1e-10):
        raise SystemExit(f"bench.py: fused and split paths disagree on the pre-flight check ({errs}); the measurement is void")
    return {"what": f"fused vs split path, 256x{np_}, 2 steps: relative L2 of f (e-, p+) and J", "rel_l2": errs}


def cpu_baseline(budget_steps=6):
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    nx, np_ = REF_SAMPLE
    cores = os.cpu_count() or 1
    if not os.path.exists(harness):
        return None
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([harness, "/dev/null", str(nx), str(np_), "1", str(DENSITY), str(budget_steps), "time_only=1"],
                         env=env, stdout=subprocess.PIPE, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("ORACLE_TIMING")][-1]
    kv = dict(x.split("=") for x in line.split()[1:])
    return {"value": float(kv["cell_updates_per_s_per_stage"]), "unit": "cell-updates/s", "cores": cores, "kind": "reference",
            "sample": f"{REF_SAMPLE_TEXT}; {budget_steps} steps of SolverManager::Advance from t=3T, OMP_NUM_THREADS={cores}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=None)
    ap.add_argument("--scaling", default="strong")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-fields-phase", action="store_true")
    ap.add_argument("--no-self-check", action="store_true", help="skip the pre-flight fused-vs-split check (profiling runs: keeps its launches out of the capture window)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    if args.workload in HOST_WORKLOADS:
        if world > 1:
            raise SystemExit("bench.py: the AMR workloads are single-GPU (whole patches are not sharded; SURVEY.md §8(e))")
        print(json.dumps(host_workload_line(args, "ours")))
        return

    import hashlib
    import numpy as np
    import torch
    import veritas_b200 as vb
    from veritas_b200 import solver as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (veritas_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    wl = args.workload or ("c3" if n_gpus == 1 else "c5")
    nx, np_ = WORKLOADS[wl]
    scaling = "weak" if n_gpus == 1 else args.scaling
    if n_gpus > 1 and args.scaling == "weak":
        nx = WORKLOADS["c3"][0] * n_gpus
        wl = f"c3 per GPU ({nx}x{np_})"

    self_check = fused_vs_split_self_check(vb, S, np_, local_rank) if (rank == 0 and not args.no_self_check) else None

    run = vb.LaserPlasmaRun(nx, np_, density=DENSITY, device=local_rank, slab=(rank, n_gpus) if n_gpus > 1 else None,
                            graph=not args.no_graph)
    ctx = run.ctx
    if n_gpus > 1:
        from veritas_b200.parallel import broadcast_unique_id
        ctx.call("vrt_comm_init", broadcast_unique_id(dist, run.L, rank, device="cuda"), rank, n_gpus)
    run.init_device()
    fields_steps = 0
    if not args.skip_fields_phase:
        fields_steps = run.run_fields_phase()          # veritas.cpp:139-144: the laser enters the box while the plasma is frozen
    else:
        run.time = 3 * run.T
    ctx.sync()
    cells = run.cells()                 # global, both species
    stream = torch.cuda.ExternalStream(ctx.L.vrt_stream(ctx.h), device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def species_charge():
        """sum over x of each species' charge density (= q N / dx): conserved to round-off by the flux form with closed walls"""
        ctx.moments()
        return [float(np.sum(ctx.get_1d(S.CHARGES0 + s))) for s in range(2)]

    q0 = species_charge()

    # ---- warm-up --------------------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    dt = run.calculate_dt()
    for _ in range(args.warmup):
        run.advance(dt)
    ctx.sync()

    # ---- device-resident throughput: K steps, CUDA events on the launching stream ----------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mark_lo = sampler.mark()
    e0.record(stream)
    for _ in range(args.steps):
        run.advance(dt)
    e1.record(stream)
    barrier()
    mark_hi = sampler.mark()
    ms = e0.elapsed_time(e1)
    launches = ctx.last_step_launches() * args.steps
    if dist is not None:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = cells * 6.0 * args.steps / (ms * 1e-3)

    # ---- end to end through the host-facing API ----------------------------------------------------------------------
    # per step: CalculateDt (device reduction + 8 B D2H), laser values + dt (13 doubles H2D as launch parameters),
    # Advance, and the 1-D arrays fileOutput would write (charge x2, PHI, E_x, a^2, Ey, Ez, By, Bz) D2H on rank 0.
    d2h = 8 + 8 * (2 * nx + nx + nx + (nx + 1) + 4 * (nx + 4))
    h2d = 13 * 8
    # host buffers of the 1-D outputs: pinned, as the contract asks for the host side of the timed copies
    pinned = lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    hb = {k: pinned(nx + 1 if k == S.A_SQUARED else nx) for k in (S.CHARGES0, S.CHARGES0 + 1, S.PHI, S.EFIELD, S.A_SQUARED)}
    hf = {w: pinned(nx + 4) for w in (S.EY, S.EZ, S.BY, S.BZ)}
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dte = run.calculate_dt()
        run.advance(dte)
        if rank == 0:        # the 1-D arrays are replicated on every rank; one process writes the output files
            for k, buf in hb.items():
                ctx.get_1d(k, buf)
            for w, buf in hf.items():
                ctx.download_field(w, 0, buf)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop(mark_lo, mark_hi)
    if dist is not None:
        ts = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e_s = float(ts.item())
    e2e_value = cells * 6.0 * args.steps / e2e_s

    # ---- roofline of the dominant kernel (fused Vlasov stage), per-launch CUDA events ---------------------------------
    cells_loc = (nx // n_gpus) * np_
    tot_bytes, tot_ms, per_stage = 0.0, 0.0, []
    reps = 2
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(12)] for _ in range(reps)]
    other = {"moments": [], "poisson": [], "field_stage": []}

    def timed(name, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        other[name].append((a, b))

    for r in range(reps):
        lasers, tnew = run.stage_lasers(dt)
        k = 0
        for i in range(6):
            timed("moments", ctx.moments); timed("poisson", ctx.poisson)
            for s in range(2):
                a, b = ev[r][k]; k += 1
                a.record(stream)
                ctx.vlasov_stage(s, dt, i)
                b.record(stream)
            timed("field_stage", lambda: ctx.field_stage(i, dt, lasers[2 * i], lasers[2 * i + 1]))
        run.time = tnew
    barrier()
    breakdown = {k: round(sum(a.elapsed_time(b) for a, b in v) / reps, 3) for k, v in other.items()}
    for i in range(6):
        msl = [ev[r][2 * i + s][0].elapsed_time(ev[r][2 * i + s][1]) for r in range(reps) for s in range(2)]
        avg = sum(msl) / len(msl)
        per_stage.append(round(stage_bytes(cells_loc, i) / (avg * 1e-3) / 1e9, 1))
        tot_bytes += stage_bytes(cells_loc, i); tot_ms += avg
    achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": None, "kernel": "k_fused_stage<S> (76 B/cell/stage algorithmic, averaged over the 6 stages; the kernel moves 13.3 B/cell/stage more: +16 for the stored stage-0 low-order fluxes, -16 at stage 5 which skips the pair with zero tableau weight)",
                "per_stage_GBps": per_stage, "peak_source": peak_src,
                "stage_cell_updates_per_s": round(6 * cells_loc / (tot_ms * 1e-3), 1)}
    breakdown["fused_stage"] = round(2 * tot_ms, 3)     # both species
    # second roof (north_star: "the slower of fp64 peak and HBM bandwidth"): fp64-pipe instructions the kernel issues per cell
    # (interior x-loop of the shipped SASS, tools/sass_count.py -> profiles/fused_sass_counts.json, times the recomputed strip /
    # chunk halo) against the fp64 issue peak measured in this session (tools/fp64_peak, a DFMA micro-benchmark)
    try:
        sc = json.load(open(os.path.join(ROOT, "profiles", "fused_sass_counts.json")))
        plan = ctx.fused_plan(0)
        halo = (plan["W"] / (plan["W"] - 6.0)) * ((nx // n_gpus) / plan["chunks"] + 6.0) / ((nx // n_gpus) / plan["chunks"])
        per_cell = [sc["stages"][f"S{i}"]["fp64_per_column"] * halo for i in range(6)]
        fp64_peak, fp64_src = measure_fp64_peak()
        stage_ms = [sum(ev[r][2 * i + s][0].elapsed_time(ev[r][2 * i + s][1]) for r in range(reps) for s in range(2)) / (2 * reps) for i in range(6)]
        instr = sum(per_cell) * cells_loc
        fp64 = {"instr_per_cell": round(sum(per_cell) / 6, 1), "halo_factor": round(halo, 4), "peak_instr_per_s": fp64_peak, "peak_source": fp64_src,
                "achieved_instr_per_s": instr / (tot_ms * 1e-3), "frac": round(instr / (tot_ms * 1e-3) / fp64_peak, 4),
                "per_stage_frac": [round(per_cell[i] * cells_loc / (stage_ms[i] * 1e-3) / fp64_peak, 3) for i in range(6)],
                "sass_version": sc.get("kernel_version"),
                "bound_cell_updates_per_s": fp64_peak / (sum(per_cell) / 6)}
        roofline["fp64"] = fp64
        hbm_bound = peak * 1e9 / 76.0
        roofline["hbm_bound_cell_updates_per_s"] = hbm_bound
        roofline["binding"] = "fp64" if fp64["bound_cell_updates_per_s"] < hbm_bound else "hbm"
        roofline["frac_of_binding"] = round(roofline["stage_cell_updates_per_s"] / min(hbm_bound, fp64["bound_cell_updates_per_s"]), 4)
    except Exception as e:      # the roof is context, never a reason to lose the line
        roofline["fp64"] = {"unavailable": repr(e)}
    # DRAM traffic of the stage kernel from the committed ncu --set full capture of the shipped kernel version (same workload
    # only): `traffic` = one launch of stage 3 (the median stage), next to the algorithmic bytes and the bytes the kernel is
    # designed to move; the other captured stages are listed beside it
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json")))
        if wl == "c3" and n_gpus == 1:
            if tr.get("kernel_version") == FUSED_KERNEL_VERSION:
                st = tr["stages"]
                head = "3" if "3" in st else sorted(st)[0]
                roofline["traffic"] = st[head]["dram_bytes_read"] + st[head]["dram_bytes_write"]
                roofline["traffic_unit"] = f"bytes per launch of stage {head} (ncu dram__bytes_read.sum + dram__bytes_write.sum)"
                roofline["traffic_source"] = tr["source"]
                roofline["traffic_algorithmic"] = float(stage_bytes(cells_loc, int(head)))
                roofline["traffic_moved_by_design"] = float(stage_bytes_moved(cells_loc, int(head)))
                roofline["traffic_per_stage"] = {k: {"dram": v["dram_bytes_read"] + v["dram_bytes_write"], "algorithmic": float(stage_bytes(cells_loc, int(k))),
                                                     "by_design": float(stage_bytes_moved(cells_loc, int(k)))} for k, v in sorted(st.items())}
            else:   # no ncu --set full capture of this kernel version yet: say so instead of passing an older one off
                roofline["traffic_note"] = f"not captured for {FUSED_KERNEL_VERSION} (profiles/fused_traffic.json holds {tr.get('kernel_version')})"
    except Exception:
        pass
    run_steps = args.warmup + 2 * args.steps + reps

    # ---- sanity of the state the numbers were measured on ---------------------------------------------------------------
    q1 = species_charge()
    fields_ok = all(bool(np.all(np.isfinite(ctx.get_1d(w)))) for w in (S.CHARGE, S.J, S.A_SQUARED, S.EFIELD))
    fields_ok = fields_ok and bool(np.all(np.isfinite(ctx.download_field(S.EY, 0))))
    drift = [abs(b - a) / abs(a) for a, b in zip(q0, q1)]
    a2max = float(np.max(ctx.get_1d(S.A_SQUARED)))
    if not fields_ok or not all(np.isfinite(drift)) or max(drift) > 1e-9:
        raise SystemExit(f"bench.py: state is not finite / particle number drifted ({drift}); the measurement is void")

    # fingerprint of the state the run ends in (rho, J, PHI, E_y, a^2 on rank 0): runs of the same workload and step counts on
    # 2, 4 and 8 GPUs must print the same hash — the x-slab decomposition is bitwise transparent
    hsh = hashlib.sha256()
    for arr in (ctx.get_1d(S.CHARGE), ctx.get_1d(S.J), ctx.get_1d(S.PHI), ctx.download_field(S.EY, 0), ctx.get_1d(S.A_SQUARED)):
        hsh.update(np.ascontiguousarray(arr).tobytes())
    state_hash = hsh.hexdigest()[:32]

    if rank == 0:
        cb = None if (args.no_cpu_baseline or n_gpus > 1) else cpu_baseline()     # the CPU baseline is timed at N = 1 only
        out = {
            "metric": "phase-space cell-updates/s per RK stage", "value": value, "unit": "cell-updates/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wl}: uniform {nx}x{np_} x-p mesh, single level, 2 species (e-, p+), laser-plasma case n={DENSITY} N_c, "
                                   f"t>=3T; inputs larger than L2 ({cells * 8 / 2**30:.1f} GiB per f plane pair)",
                       "parallelism": f"x-slab x{n_gpus}" if n_gpus > 1 else "single GPU", "cuda_graph": (not args.no_graph) and (n_gpus == 1 or os.environ.get("VRT_MULTI_GRAPH", "1") != "0")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "CalculateDt + Advance + 1-D output arrays per step through the C ABI with host buffers; f stays resident as in the reference"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "breakdown_ms_per_step": breakdown,
            "checks": {"finite": fields_ok, "particle_number_rel_drift": drift, "max_a_squared": a2max,
                       "fields_phase_steps": fields_steps, "steps_run": run_steps, "state_hash": state_hash,
                       "self_check": self_check},
            "cpu_baseline": cb,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
