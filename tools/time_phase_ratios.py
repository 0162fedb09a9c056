"""Timing of update_phase_ratios! (centre, vertex, faces, midpoints) -- developer tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import justpic.jl_b200 as J
from tests.problems import make_grids, stream_velocity, cfl_dt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gr = make_grids(n, 3, True)
p = J.init_particles(J.CUDABackend, 24, 48, 12, *gr.grid_vel, seed=42)
Vn = stream_velocity(gr); V = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in Vn]
dt = cfl_dt(gr, Vn, 0.5)
ph, = J.init_cell_arrays(p, 1)
ph.copy_(torch.where(p.index > 0, 1.0 + (p.coords[0] < p.coords[2]).double(), torch.zeros_like(ph)))
for _ in range(5):
    J.advection(p, J.RungeKutta2(), V, dt); J.move_particles(p, (ph,))
pr = J.PhaseRatios(J.CUDABackend, 2, gr.n)
calls = [("center", lambda: J.phase_ratios_center(pr, p, ph)), ("vertex", lambda: J.phase_ratios_vertex(pr, p, ph)),
         ("Vx", lambda: J.phase_ratios_face(pr.Vx, p, ph, "x")), ("Vy", lambda: J.phase_ratios_face(pr.Vy, p, ph, "y")),
         ("Vz", lambda: J.phase_ratios_face(pr.Vz, p, ph, "z")), ("xy", lambda: J.phase_ratios_midpoint(pr.xy, p, ph, "xy")),
         ("yz", lambda: J.phase_ratios_midpoint(pr.yz, p, ph, "yz")), ("xz", lambda: J.phase_ratios_midpoint(pr.xz, p, ph, "xz"))]
tot = 0.0
for name, fn in calls:
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3; tot += ms
    print(f"{name:7s} {ms:8.3f} ms")
print(f"update_phase_ratios total {tot:8.3f} ms at {n}^3")

for mode in ("literal", "fused"):
    J.update_phase_ratios(pr, p, ph, mode=mode); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): J.update_phase_ratios(pr, p, ph, mode=mode)
    e1.record(); torch.cuda.synchronize()
    print(f"update_phase_ratios[{mode}] {e0.elapsed_time(e1) / 3:8.3f} ms at {n}^3")
