"""Per-step, per-phase CUDA-event timings of the headline step (developer tool).
Usage: python tools/time_phases.py [--cells 128] [--steps 12] [--fields 3]
Set JUSTPIC_LIB=<path to another build> to A/B two builds on identical inputs."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import justpic.jl_b200 as J
from tests.problems import make_grids, stream_velocity, cfl_dt, vertex_field_linear

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=128)
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--fields", type=int, default=3)
ap.add_argument("--ndim", type=int, default=3)
ap.add_argument("--method", default="rk2")
ap.add_argument("--exact", action="store_true", help="Julia range() grids (exactly affine centres)")
ap.add_argument("--affine", type=int, default=1)
ap.add_argument("--policy", default="reference", help="move slot policy: reference | compact")
ap.add_argument("--classify", type=int, default=0, help="1: advection -> move hand-off (JP_OPT_ADVECT_CLASSIFY)")
ap.add_argument("--interp", type=int, default=0, help="1: move -> interpolation hand-off (JP_OPT_MOVE_INTERP)")
a = ap.parse_args()
gr = make_grids(a.cells, a.ndim, True, exact=a.exact)
p = J.init_particles(J.CUDABackend, 24, 48, 12, *gr.grid_vel, seed=42)
Vn = stream_velocity(gr)
V = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in Vn]
dt = cfl_dt(gr, Vn, 0.5)
T = torch.from_numpy(np.ascontiguousarray(vertex_field_linear(gr))).cuda()
fields = J.init_cell_arrays(p, a.fields)
J.grid2particle(fields[0], T, p)
if a.fields > 1:
    fields[1].copy_(torch.where(p.index > 0, 1.0 + (p.coords[0] < p.coords[-1]).double(), torch.zeros_like(fields[0])))
pr = J.PhaseRatios(J.CUDABackend, 2, gr.n)
m = {"rk2": J.RungeKutta2(), "rk4": J.RungeKutta4(), "euler": J.Euler()}[a.method]
if a.interp:
    J.move_interp_handoff(p, Fp=fields[0], phases=fields[1] if a.fields > 1 else None, nphases=2)
J.profile_move(p, True)
names = ["advect", "move", "p2g", "phase"]
rows = []
for it in range(a.steps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record(); J.advection(p, m, V, dt, affine=bool(a.affine), classify=(True if a.classify else None))
    ev[1].record(); J.move_particles(p, tuple(fields), policy=a.policy)
    ev[2].record(); J.particle2grid(T, fields[0], p)
    ev[3].record()
    if a.fields > 1: J.phase_ratios_center(pr, p, fields[1])
    ev[4].record(); torch.cuda.synchronize()
    rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
r = np.array(rows)
print("move stages (mean ms over all steps)", {k: round(v, 3) for k, v in J.read_move_profile(p).items()}, "interp hand-off used", J.last_interp_handoff(p))
print("classify", a.classify, J.last_move_classify(p) if a.classify else "-", "policy", a.policy, "lib", os.environ.get("JUSTPIC_LIB", "default"), "affine", J.advect_affine_level(p) if hasattr(J, "advect_affine_level") else None, "cells", a.cells, "live", int(p.index.sum()))
for i, n in enumerate(names):
    print(f"{n:8s}", " ".join(f"{x:7.3f}" for x in r[:, i]), f"| mean(last half) {r[len(r)//2:, i].mean():7.3f}")
print("checksum", float(torch.nan_to_num(p.coords[0]).sum()), float(torch.nan_to_num(fields[0]).sum()), int(p.index.sum()))
