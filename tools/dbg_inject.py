import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import justpic.jl_b200 as J
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gv = bench.local_grids(n)
p = J.init_particles(J.CUDABackend, 24, 48, 12, *gv, seed=42)
V = [torch.from_numpy(v).cuda() for v in bench.stream_velocity_np(gv)]
dt = 0.5 * min(p.di.vertex[0] / 250, p.di.vertex[2] / 250)
pT, ph, st = J.init_cell_arrays(p, 3)
T = torch.zeros((n + 1,) * 3, dtype=torch.float64, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for it in range(8):
    ev[0].record(); J.advection(p, J.RungeKutta2(), V, dt)
    ev[1].record(); J.move_particles(p, (pT, ph, st))
    ev[2].record(); J.inject_particles(p, (pT, ph, st))
    ev[3].record(); J.grid2particle(pT, T, p)
    ev[4].record(); J.particle2grid(T, pT, p)
    ev[5].record(); torch.cuda.synchronize()
    print(it, "advect %.2f move %.2f inject %.2f g2p %.2f p2g %.2f" % tuple(ev[i].elapsed_time(ev[i + 1]) for i in range(5)),
          "injected", J.inject_stats(p), "live", int(p.index.sum()), flush=True)
