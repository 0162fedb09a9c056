"""advection_LinP! / advection_MQS! timing (developer tool): python tools/time_interpolants.py [--cells 128]
JP_ADVECT_HI_GLOBAL=1 forces the thread-per-cell global-memory kernel instead of the tiled one."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import justpic.jl_b200 as J
from tests.problems import make_grids, stream_velocity, cfl_dt
ap = argparse.ArgumentParser(); ap.add_argument("--cells", type=int, default=128); a = ap.parse_args()
gr = make_grids(a.cells, 3, True)
p = J.init_particles(J.CUDABackend, 24, 48, 12, *gr.grid_vel, seed=42)
Vn = stream_velocity(gr); V = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in Vn]
dt = cfl_dt(gr, Vn, 0.5)
for it in range(3):
    J.advection(p, J.RungeKutta2(), V, dt); J.move_particles(p)
for name, fn in (("linear", J.advection), ("LinP", J.advection_LinP), ("MQS", J.advection_MQS)):
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(p, J.RungeKutta2(), V, dt); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); J.move_particles(p)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(p, J.RungeKutta2(), V, 0.1 * dt); e0.record(); fn(p, J.RungeKutta2(), V, 0.1 * dt); e1.record(); torch.cuda.synchronize()
    J.move_particles(p)
    print(f"{name:7s} RK2 second advection! in a row (no move_particles! between: first interpolation re-centres): {e0.elapsed_time(e1):7.3f} ms")
    print(f"{name:7s} RK2 {a.cells}^3: " + " ".join(f"{t:7.3f}" for t in ts) + " ms", "(global-memory kernel)" if os.environ.get("JP_ADVECT_HI_GLOBAL") and name != "linear" else "")
print("checksum", float(torch.nan_to_num(p.coords[0]).sum()), int(p.index.sum()))
