#!/usr/bin/env python
"""bench.py -- headline benchmark of the JustPIC hot path on B200.

Metric (BASELINE.json): particle-updates/s per step, one update = one live
particle taken through advection! (RK2) + move_particles! + particle2grid!(T)
(+ phase_ratios_center! in the headline configuration).  Workload at N=1 is
BASELINE configs[3]: 3D 256^3 cells, 24 particles/cell (48 slots), 3 advected
fields (T, phase, strain), 2 phases, stream-function velocity, CFL 0.5.
N>1 (torchrun): weak scaling, one 256^3 block per GPU, block decomposition with
the halo exchange of configs[4] between advection! and move_particles!.

Prints ONE JSON line (see DESIGN.md "Measurement" for every key).
    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--cells CELLS]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PPC, SLOTS, MIN_XCELL, NFIELDS, NPHASES, CFL = 24, 48, 12, 3, 2, 0.5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", dest="n", type=int, default=256, help="cells per dimension per GPU (headline: 256)")
    ap.add_argument("--cpu-n", type=int, default=64, help="cells per dimension of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--handoff", type=int, default=1, choices=[0, 1],
                    help="1 (default): advection! also leaves move_particles!' classification words (JP_OPT_ADVECT_CLASSIFY, "
                         "bit-identical results, see include/justpic_c.h); 0: move_particles! classifies the coordinates itself")
    ap.add_argument("--overlap", type=int, default=1, choices=[0, 1],
                    help="N > 1 only. 1 (default): advection! runs as shell + interior launches and update_cell_halo! travels on a side "
                         "stream behind the interior launch (halo.advection_with_halo, bit-identical results); 0: advection!, then the exchange")
    return ap.parse_args()


# ----------------------------------------------------------------------------- workload
def local_grids(n, topo_dims=(1, 1, 1), coords=(0, 0, 0)):
    """Unit cube split into blocks of n cells with a 1-cell halo ring (overlap 2,
    as ImplicitGlobalGrid does); returns LinRange staggered grids of this block."""
    from justpic.jl_b200 import LinRange, expand_range
    xv, xc = [], []
    for d in range(3):
        nglob = topo_dims[d] * (n - 2) + 2 if topo_dims[d] > 1 else n
        dx = 1.0 / nglob
        i0 = coords[d] * (n - 2) if topo_dims[d] > 1 else 0
        v = LinRange(i0 * dx, (i0 + n) * dx, n + 1)
        xv.append(v)
        xc.append(LinRange(i0 * dx + dx / 2, (i0 + n) * dx - dx / 2, n))
    xg = [expand_range(c) for c in xc]
    return tuple(tuple(xv[d] if d == comp else xg[d] for d in range(3)) for comp in range(3))


def stream_velocity_np(grid_vel):
    V = []
    for comp in range(3):
        x = np.asarray(grid_vel[comp][0])[None, None, :]
        z = np.asarray(grid_vel[comp][2])[:, None, None]
        shape = tuple(len(grid_vel[comp][d]) for d in (2, 1, 0))
        if comp == 0:
            v = 250.0 * np.sin(math.pi * x) * np.cos(math.pi * z)
        elif comp == 2:
            v = -250.0 * np.cos(math.pi * x) * np.sin(math.pi * z)
        else:
            v = np.zeros((1, 1, 1))
        V.append(np.ascontiguousarray(np.broadcast_to(v, shape), dtype=np.float64))
    return V


def state_checksum(p, pT):
    """Three finite numbers that pin rank 0's final particle state (NaN / Inf slots count as 0)."""
    try:
        import torch
        fin = lambda t: float(torch.nan_to_num(t, nan=0.0, posinf=0.0, neginf=0.0).sum().item())
        return [fin(p.coords[0]), fin(pT), int(p.index.sum().item())]
    except Exception as e:           # never lose the bench line over a diagnostic
        return [repr(e)]


def algorithmic_bytes(f_mig):
    """SURVEY.md section 8(d), per live particle, N=3, S=48, ppc=24, F=3 fields."""
    adv = 16 * 3 + SLOTS / PPC + 8 * 3 / PPC
    mov = 8 * 3 + SLOTS / PPC + f_mig * (2 * 8 * 3 + 3 * 8 * NFIELDS + 2)
    p2g = 8 * 3 + SLOTS / PPC + 8 + 8 / PPC
    phr = 8 * 3 + 8 + SLOTS / PPC + 8 * NPHASES / PPC
    return {"advect": adv, "move": mov, "p2g": p2g, "phase_ratios": phr}


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "power_w_max": max((float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()), default=None),
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ----------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_reference_run(n, steps, warmup, threads):
    """The reference's CPU path restated (oracle/) on a bounded sample of the
    same workload: 3D n^3 cells, same ppc / fields / kernels, all host threads."""
    from oracle.oracle import Oracle
    from justpic.jl_b200 import LinRange  # noqa: F401  (grid helper only)
    gv = local_grids(n)
    xi_vel = tuple(tuple(np.asarray(x, dtype=np.float64) for x in g) for g in gv)
    xvi = tuple(xi_vel[i][i] for i in range(3))
    xci = (xi_vel[1][0][1:-1].copy(), xi_vel[0][1][1:-1].copy(), xi_vel[0][2][1:-1].copy())
    o = Oracle(xvi, xci, xi_vel, SLOTS, True)
    Oracle.set_threads(threads)
    coords, index = o.init_particles(PPC, 42)
    V = stream_velocity_np(gv)
    dt = CFL * min((xvi[0][1] - xvi[0][0]) / np.abs(V[0]).max(), (xvi[2][1] - xvi[2][0]) / np.abs(V[2]).max())
    T = np.ascontiguousarray(np.broadcast_to(xvi[2][:, None, None], (n + 1, n + 1, n + 1)))
    pT = np.zeros_like(coords[0]); o.grid2particle(coords, index, pT, T)
    ph = np.where(index > 0, 1.0 + (coords[0] < coords[2]), 0.0)
    strain = np.zeros_like(pT)
    ratios = np.zeros((NPHASES, n, n, n))
    F = np.empty_like(T)
    times, updates = [], 0
    for it in range(warmup + steps):
        live = int(index.sum())
        t0 = time.perf_counter()
        o.advect(coords, index, 1, 0.5, V, dt)
        o.move(coords, index, [pT, ph, strain])
        o.particle2grid(coords, index, F, pT)
        o.phase_ratios_center(coords, ratios, ph, NPHASES)
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0); updates += live
    Oracle.set_threads(1)
    total = sum(times)
    return updates / total, total / len(times) * 1e3, f"3D {n}^3 cells, {PPC} ppc, RK2 advect+move+p2g+phase_ratios, {len(times)} steps"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import Oracle
    threads = Oracle.max_threads()
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    v, ms, sample = cpu_reference_run(args.cpu_n, steps, warmup, threads)
    line = {
        "impl": "reference", "metric": "particle-updates/s per step (advect+move+p2g)", "value": v, "unit": "particle-updates/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.gpus), "note": "JustPIC.CPU cannot run here (no Julia); CPU restatement (oracle/, OpenMP over same-colour cells) timed on a bounded sample"},
        "cpu_baseline": {"value": v, "unit": "particle-updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(n, gpus):
    return (f"3D {n}^3 cells/GPU x {gpus} GPU, {PPC} ppc ({SLOTS} slots), RK2(0.5) advection! + "
            f"{'update_cell_halo! + ' if gpus > 1 else ''}move_particles!({NFIELDS} fields) + particle2grid!(T) + "
            f"phase_ratios_center!({NPHASES} phases), stream-function velocity, CFL {CFL}")


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import justpic.jl_b200 as J
    from justpic.jl_b200.halo import CartesianTopology, update_cell_halo, advection_with_halo, join_halo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the JustPIC hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        opts = None
        if args.overlap:
            # the exchange runs behind the interior advection launch: NCCL's copy kernels must win SM slots as advection
            # CTAs retire (high-priority stream) and fit into one such slot (256 threads per channel instead of 640)
            os.environ.setdefault("NCCL_NTHREADS", "256")
            opts = dist.ProcessGroupNCCL.Options()
            opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    topo = CartesianTopology.create(world, 3, rank)
    n = args.n
    gv = local_grids(n, topo.dims, topo.coords())
    p = J.init_particles(J.CUDABackend, PPC, SLOTS, MIN_XCELL, *gv, seed=42 + rank, device=dev)
    V_host = [torch.from_numpy(v).pin_memory() for v in stream_velocity_np(gv)]
    V = [v.to(dev, non_blocking=True) for v in V_host]
    vmax = torch.tensor([float(np.abs(V_host[0].numpy()).max()), float(np.abs(V_host[2].numpy()).max())], device=dev)
    if world > 1:
        dist.all_reduce(vmax, op=dist.ReduceOp.MAX)     # dt = MPI.Allreduce(max) in the reference script
    dx = p.di.vertex[0]
    dt = CFL * min(dx / float(vmax[0]), p.di.vertex[2] / float(vmax[1]))
    zv = torch.from_numpy(np.asarray(p.xvi[2])).to(dev)
    T = zv[:, None, None].expand(n + 1, n + 1, n + 1).contiguous()
    T_host = torch.empty(T.shape, dtype=T.dtype).pin_memory()
    pT, ph, strain = J.init_cell_arrays(p, NFIELDS)
    J.grid2particle(pT, T, p)
    ph.copy_(torch.where(p.index > 0, 1.0 + (p.coords[0] < p.coords[2]).double(), torch.zeros_like(pT)))
    pr = J.PhaseRatios(J.CUDABackend, NPHASES, (n, n, n), device=dev)
    fields = (pT, ph, strain)
    rk2 = J.RungeKutta2()
    halo_buffers = {}
    phases = ["advect", "halo", "move", "p2g", "phase_ratios"]

    def step(ev=None):
        def mark(i):
            if ev is not None:
                ev[i].record()
        mark(0)
        if world > 1 and args.overlap:
            # shell bricks -> [side stream: pack / NCCL / unpack] || interior bricks -> join: the "halo" phase below is what is
            # left of the exchange after the interior launch has finished (the "advect" phase holds both launches)
            advection_with_halo(p, rk2, V, dt, fields, topo, buffers=halo_buffers, classify=bool(args.handoff), join=False)
            mark(1)
            join_halo(p)
        else:
            J.advection(p, rk2, V, dt, classify=bool(args.handoff))
            mark(1)
            if world > 1:
                update_cell_halo(p, fields, topo, buffers=halo_buffers)
        mark(2)
        J.move_particles(p, fields)
        mark(3)
        J.particle2grid(T, pT, p)
        mark(4)
        J.phase_ratios_center(pr, p, ph)
        mark(5)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    live0 = int(p.index.sum().item())
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
    lives, migr = [], []
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for k in range(args.steps):
        step(evs[k])
    t_end.record()
    barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    clocks = sampler.stop() if sampler else None
    # particle count / migrant fraction are read AFTER the timed region (one extra step, untimed)
    live = int(p.index.sum().item())
    step()
    moved, dropped, deleted = J.move_stats(p)
    move_path = J.last_move_path(p)
    move_classify = J.last_move_classify(p)
    f_mig = (moved + dropped + deleted) / max(live, 1)
    # particles processed per step ~ live count (changes by drops only); use mean of start/end
    updates = 0.5 * (live0 + live) * args.steps
    tmax = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    tot_updates = torch.tensor([updates], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_updates, op=dist.ReduceOp.SUM)
    elapsed_ms = float(tmax.item())
    value = float(tot_updates.item()) / (elapsed_ms * 1e-3)
    per_phase = {ph_: float(np.mean([evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(args.steps)])) for i, ph_ in enumerate(phases)}

    # ---- e2e: same step through the public API with HOST buffers.  Every step the velocity field
    # (the Stokes solver's output in the reference's time loop) comes from pinned host memory and
    # the grid field T + the live-particle count go back to the host.  Copies run on a side stream:
    # V for step k+1 is uploaded into the second device buffer while step k computes, T of step k
    # is downloaded while step k+1 computes (double buffering; every byte is still copied every step,
    # inside the timed region, and the host waits for T before the loop ends).
    e2e = None
    if not args.no_e2e:
        h2d = sum(v.numel() * 8 for v in V_host)
        d2h = T_host.numel() * 8 + 8
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream()
        Vbuf = [V, [torch.empty_like(v) for v in V]]
        Tbuf = [T, torch.empty_like(T)]
        T_hosts = [T_host, torch.empty_like(T_host).pin_memory()]
        nlive_host = torch.zeros(2, dtype=torch.int64).pin_memory()
        up_done = [torch.cuda.Event(), torch.cuda.Event()]
        comp_done = [torch.cuda.Event(), torch.cuda.Event()]
        down_done = [torch.cuda.Event(), torch.cuda.Event()]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        esteps = max(2, args.steps)          # same step count as the device-timed region: the un-overlapped first upload / last download amortise alike
        live_a = int(p.index.sum().item())

        def upload(k):
            with torch.cuda.stream(copy_stream):
                if k >= 2:
                    copy_stream.wait_event(comp_done[k % 2])        # buffer k%2 was last read by step k-2
                for vd, vh in zip(Vbuf[k % 2], V_host):
                    vd.copy_(vh, non_blocking=True)
                up_done[k % 2].record(copy_stream)

        def e2e_step(k):
            nonlocal V, T
            main_stream.wait_event(up_done[k % 2])
            if k >= 2:
                main_stream.wait_event(down_done[k % 2])                 # T buffer k%2 still being downloaded (step k-2)
            V, T = Vbuf[k % 2], Tbuf[k % 2]
            step()
            nl = p.index.sum()
            comp_done[k % 2].record(main_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(comp_done[k % 2])
                T_hosts[k % 2].copy_(T, non_blocking=True)
                nlive_host[k % 2:k % 2 + 1].copy_(nl.reshape(1), non_blocking=True)
                down_done[k % 2].record(copy_stream)

        barrier()
        e0.record()
        upload(0)
        for k in range(esteps):
            if k + 1 < esteps:
                upload(k + 1)
            e2e_step(k)
            if k >= 1:
                down_done[(k - 1) % 2].synchronize()              # host consumes T / live count of step k-1
                _ = float(T_hosts[(k - 1) % 2][0, 0, 0]) + int(nlive_host[(k - 1) % 2])
        down_done[(esteps - 1) % 2].synchronize()
        _ = float(T_hosts[(esteps - 1) % 2][0, 0, 0]) + int(nlive_host[(esteps - 1) % 2])
        main_stream.wait_stream(copy_stream)
        e1.record()
        barrier()
        V, T = Vbuf[0], Tbuf[0]
        live_b = int(p.index.sum().item())
        ems = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        eupd = torch.tensor([0.5 * (live_a + live_b) * esteps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            dist.all_reduce(eupd, op=dist.ReduceOp.SUM)
        e2e = {"value": float(eupd.item()) / (float(ems.item()) * 1e-3), "unit": "particle-updates/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": float(ems.item()) / esteps, "steps": esteps,
               "what": "per step: velocity field V (3 staggered arrays) H2D from pinned memory, hot path through the public API, "
                       "grid field T + live count D2H; copies double-buffered on a side stream"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        ab = algorithmic_bytes(f_mig)
        nlive_mean = 0.5 * (live0 + live)
        kernel_gbs = {k: ab[k] * nlive_mean / (per_phase[k] * 1e-3) / 1e9 for k in ab}
        step_bytes = sum(ab.values()) * nlive_mean
        step_ms = elapsed_ms / args.steps
        traffic = None
        try:   # dram__bytes_read+write per launch of the same kernel/config from the committed ncu --set full capture
            tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())
            traffic = tj.get(f"k_advect_tile@{n}")
        except Exception:
            pass
        line = {
            "metric": "particle-updates/s per step (advect+move+p2g)", "value": value, "unit": "particle-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n, world), "cells_per_gpu": n ** 3, "live_particles_per_gpu": int(nlive_mean),
                       "migrant_fraction": round(f_mig, 4), "move_path": move_path, "dropped_per_step": dropped,
                       "advect_move_handoff": bool(args.handoff), "move_classify": move_classify,
                       "halo_overlap": bool(world > 1 and args.overlap),
                       # rank 0's final state in three numbers (compare two runs, e.g. --overlap 0 / 1: must be identical)
                       "state_checksum": state_checksum(p, pT),
                       "p2g_mode": J.api.P2G_MODE, "l2": "inputs (39 GB/GPU) far larger than L2, no flush needed",
                       "topology": list(topo.dims)},
            # per step: advect 1; move (plan path) classify 1 (hand-off: 0, + 1 per halo plane rewritten) + plan 27 +
            # finalize 1 + scan 2 + set 1 + gather 1 + scatter 1; p2g 2 (cell + node); phase ratios 1;
            # halo: 2 pack + 2 unpack per decomposed dimension
            "gpu_launches": args.steps * ((2 if world > 1 and args.overlap else 1) + 33 + (0 if args.handoff else 1) + 2 + 1
                                          + ((4 + (2 if args.handoff else 0)) * sum(1 for d in topo.dims if d > 1) if world > 1 else 0)),
            "phase_ms": per_phase,
            # dominant single kernel: k_advect_tile (the advect phase is exactly one launch, so its
            # CUDA-event time is the kernel's duration); the move phase is longer but spans 34 launches
            "roofline": {"bound": "hbm", "kernel": "k_advect_tile<3,RK2,uniform>",
                         "achieved": kernel_gbs["advect"], "peak": peak, "unit": "GB/s", "frac": kernel_gbs["advect"] / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": ab, "per_phase_GBps": kernel_gbs,
                         "per_phase_frac": {k: v / peak for k, v in kernel_gbs.items()},
                         "step_GBps": step_bytes / (step_ms * 1e-3) / 1e9,
                         "step_frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak,
                         "note": "advect is bound by instruction issue and shared-memory (LDS) latency, not HBM; with the hand-off it also "
                                 "does move_particles!' classification (its 26 B/particle scan is no longer read) but is still "
                                 "credited with the 51 B/particle of advection! only; see DESIGN.md 4.1-4.2"},
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            from oracle.oracle import Oracle
            threads = Oracle.max_threads()
            v, ms, sample = cpu_reference_run(args.cpu_n, 2, 1, threads)
            line["cpu_baseline"] = {"value": v, "unit": "particle-updates/s", "cores": threads, "kind": "port", "sample": sample,
                                    "ms_per_step": ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
