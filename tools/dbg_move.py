import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import justpic.jl_b200 as J
from justpic.jl_b200.api import last_move_reasons
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gv = bench.local_grids(n)
p = J.init_particles(J.CUDABackend, 24, 48, 12, *gv, seed=42)
V = [torch.from_numpy(v).cuda() for v in bench.stream_velocity_np(gv)]
dt = 0.5 * min(p.di.vertex[0] / 250, p.di.vertex[2] / 250)
pT, ph, st = J.init_cell_arrays(p, 3)
for it in range(5):
    J.advection(p, J.RungeKutta2(), V, dt)
    torch.cuda.synchronize()
    J.move_particles(p, (pT, ph, st))
    print(it, J.last_move_path(p), last_move_reasons(p), J.move_stats(p), flush=True)
