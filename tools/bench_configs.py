"""Per-phase timings of the other BASELINE.json configurations (parity-test cases, not bench
lines): cfg1 2-D 256^2 RK2, cfg2 2-D 512^2 RK4 + inject (rotation field), cfg3 3-D 128^3 RK2.
Usage: python tools/bench_configs.py [cfg1|cfg2|cfg3 ...]"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import justpic.jl_b200 as J
from tests.problems import make_grids, stream_velocity, rotation_velocity, cfl_dt, vertex_field_linear

CFG = {"cfg1": dict(ndim=2, n=256, method="rk2", cfl=0.75, inject=False, vel="stream"),
       "cfg2": dict(ndim=2, n=512, method="rk4", cfl=0.75, inject=True, vel="rotation"),   # reference script: dt=200 at 256 vertices; at 512 cells that is > 1 cell/step
       "cfg3": dict(ndim=3, n=128, method="rk2", cfl=0.5, inject=False, vel="stream")}

def run(name, steps=20, warmup=3):
    c = CFG[name]
    gr = make_grids(c["n"], c["ndim"], True)
    p = J.init_particles(J.CUDABackend, 24, 48, 12, *gr.grid_vel, seed=42)
    Vn = stream_velocity(gr) if c["vel"] == "stream" else rotation_velocity(gr)
    V = [torch.from_numpy(v).cuda() for v in Vn]
    dt = cfl_dt(gr, Vn, c["cfl"]) if c["cfl"] else 200.0
    T = torch.from_numpy(vertex_field_linear(gr)).cuda()
    pT, = J.init_cell_arrays(p, 1)
    J.grid2particle(pT, T, p)
    m = J.RungeKutta2() if c["method"] == "rk2" else J.RungeKutta4()
    names = ["advect", "move", "inject", "p2g", "g2p"]
    acc = {k: 0.0 for k in names}; tot = 0.0; upd = 0
    for it in range(warmup + steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        live = int(p.index.sum())
        ev[0].record(); J.advection(p, m, V, dt)
        ev[1].record(); J.move_particles(p, (pT,))
        ev[2].record()
        if c["inject"]: J.inject_particles(p, (pT,))
        ev[3].record(); J.particle2grid(T, pT, p)
        ev[4].record(); J.grid2particle(pT, T, p)
        ev[5].record(); torch.cuda.synchronize()
        if it >= warmup:
            for i, k in enumerate(names): acc[k] += ev[i].elapsed_time(ev[i + 1])
            tot += ev[0].elapsed_time(ev[5]); upd += live
    print(json.dumps({"config": name, **c, "live_particles": live, "ms_per_step": tot / steps,
                      "particle_updates_per_s": upd / (tot * 1e-3), "phase_ms": {k: v / steps for k, v in acc.items()},
                      "move_path": J.last_move_path(p)}), flush=True)

if __name__ == "__main__":
    for name in (sys.argv[1:] or ["cfg1", "cfg2", "cfg3"]):
        run(name)
