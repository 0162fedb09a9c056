"""Run EVERY kernel of libjustpic_sm100a.so once on a realistic state (developer / profiling tool).

    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/all python tools/all_kernels.py --cells 128

Three untimed warm-up steps bring the slot planes to their steady ~50 % occupancy; the profiled region
(cudaProfilerStart/Stop) then calls each public entry point once: the headline step with and without the
advection -> move hand-off, both particle2grid modes, the literal and the fused update_phase_ratios, inject /
inject_phase, LinP / MQS advection, FLIP interpolation, subgrid diffusion, clean, halo pack/unpack and the
Array()/CuArray() layout conversion.  profiles/kernel_table.py turns the raw page into the per-kernel table."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import justpic.jl_b200 as J
from justpic.jl_b200 import halo as H
from tests.problems import make_grids, stream_velocity, cfl_dt, vertex_field_linear, centre_field_linear

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=128)
ap.add_argument("--ndim", type=int, default=3)
a = ap.parse_args()
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
gr = make_grids(a.cells, a.ndim, True)
p = J.init_particles(J.CUDABackend, 24, 48, 12, *gr.grid_vel, seed=42)
Vn = stream_velocity(gr); V = [dev(v) for v in Vn]
dt = cfl_dt(gr, Vn, 0.5)
T = dev(vertex_field_linear(gr)); T0 = T.clone(); Tc = dev(centre_field_linear(gr))
pT, ph, strain = J.init_cell_arrays(p, 3)
J.grid2particle(pT, T, p)
ph.copy_(torch.where(p.index > 0, 1.0 + (p.coords[0] < p.coords[-1]).double(), torch.zeros_like(pT)))
pr = J.PhaseRatios(J.CUDABackend, 2, gr.n)
fields = (pT, ph, strain)
rk2 = J.RungeKutta2()
for it in range(3):                                   # warm-up: steady slot-plane occupancy
    J.advection(p, rk2, V, dt); J.move_particles(p, fields); J.inject_particles(p, fields, step=it)
torch.cuda.synchronize()
torch.cuda.profiler.start()
# --- headline step, default path
J.advection(p, rk2, V, dt, classify=False)
J.move_particles(p, fields)
J.particle2grid(T, pT, p)                             # two-pass (default)
J.phase_ratios_center(pr, p, ph)
J.inject_particles(p, fields, step=10)
# --- headline step with the advection -> move hand-off
J.advection(p, rk2, V, dt, classify=True)
J.move_particles(p, fields)
J.particle2grid(T, pT, p, mode="exact")
J.particle2grid(T, pT, p, mode="twopass_fastw")
J.inject_particles_phase(p, ph, (pT,), (T,), step=11)
# --- headline step the way bench.py runs it: both hand-offs (k_move_scatter_interp_fast, node pass, ratio copy)
J.move_interp_handoff(p, Fp=pT, phases=ph, nphases=2)
J.advection(p, rk2, V, dt, classify=True)
J.move_particles(p, fields)
J.particle2grid(T, pT, p)
J.phase_ratios_center(pr, p, ph)
# ... and with an argument order the fast path does not serve (generic fused scatter)
J.move_interp_handoff(p, Fp=pT, phases=ph, nphases=2)
J.advection(p, rk2, V, 0.5 * dt, classify=True)
J.move_particles(p, (strain, ph, pT))
J.particle2grid(T, pT, p)
J.move_interp_handoff(p, enable=False)
# --- the opt-in dense slot policy (k_move_prevacate + the plan without cursor)
J.advection(p, rk2, V, 0.5 * dt, classify=False)
J.move_particles(p, fields, policy="dense")
# --- other integrators / interpolants (a move between them: each advection then starts from the bucketed state, as in a time loop)
J.advection(p, J.Euler(), V, 0.1 * dt, classify=False); J.move_particles(p, fields)
J.advection(p, J.RungeKutta4(), V, 0.1 * dt); J.move_particles(p, fields)
J.advection_LinP(p, rk2, V, 0.1 * dt); J.move_particles(p, fields)
J.advection_MQS(p, rk2, V, 0.1 * dt)
J.move_particles(p, fields, mode="direct")            # literal sweeps (one cooperative launch)
J.clean_particles(p, None, fields)
# --- interpolations
J.grid2particle(pT, T, p)
J.grid2particle_flip(pT, None, T, T0, p, alpha=0.25)
J.centroid2particle(strain, Tc, p)
J.particle2centroid(Tc, pT, p)
sa = J.SubgridDiffusionCellArrays(p)
dTg = torch.zeros([n + 1 for n in T.shape], dtype=torch.float64, device="cuda")       # read at I + 1: one more node per dimension
dTg[tuple(slice(1, None) for _ in T.shape)] = T - T0
J.subgrid_diffusion(pT, T, dTg, sa, p, dt)
sac = J.SubgridDiffusionCellArrays(p, loc="center")
J.subgrid_diffusion_centroid(pT, Tc, torch.zeros([n + 1 for n in Tc.shape], dtype=torch.float64, device="cuda"), sac, p, dt)
# --- phase ratios on every node family
J.update_phase_ratios(pr, p, ph, mode="literal")
J.update_phase_ratios(pr, p, ph, mode="fused")
# --- halo planes and layout conversion
arrays = [*p.coords, *fields]
buf = torch.empty(H.plane_bytes(p.ncells, p.max_xcell, 0, len(arrays)), dtype=torch.uint8, device="cuda")
H._cuda_pack(p, 0, 1, arrays, buf); H._cuda_unpack(p, 0, 0, arrays, buf)
# jp_halo_exchange / jp_halo_exchange_grid: a rank that is its own neighbour (periodic) runs the library's schedule without NCCL
topo = H.CartesianTopology((1,) * a.ndim, 0, (True,) * a.ndim)
H.update_cell_halo(p, fields, topo)
H.update_halo(p, V[0], topo)
h = J.Array(pT); J.CuArray(h)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("all kernels ran; live particles", int(p.index.sum()))
