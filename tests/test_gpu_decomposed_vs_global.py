"""A block-decomposed run == the same problem on the undecomposed (global) grid in the oracle.

Set-up of scripts/temperature_advection3D_MPI.jl:27-91 (ImplicitGlobalGrid: local blocks with a 1-cell halo ring, overlap 2;
per step advection! -> update_cell_halo!(coords..., args..., index) -> move_particles!): every rank starts from its block
of the global initial state, runs the library's decomposed path, and after every step the cells it OWNS must hold exactly
the particles -- coordinates and fields bit for bit -- that the oracle's run on the global grid holds in the matching
global cells.  Grids use a power-of-two spacing so that block and global coordinates are the same doubles, and blocks
whose owned width is a multiple of 3 so that block-local and global sweep colours coincide.

Slot positions inside a cell are compared canonically (particles of a cell sorted by x): the reference's free-slot cursor
is carried from one migrant of a source cell to the next ACROSS destination cells (src/Particles/move_safe.jl:114-118), so a
migrant placed into a halo cell -- whose occupancy differs from the global run's, because the cells beyond it do not exist
on this rank -- shifts the slots of the source cell's later migrants.  ImplicitGlobalGrid runs of the reference differ from
its single-rank runs in exactly the same way; which particles are in which cell does not -- as long as no particle is
DROPPED: a migrant is dropped when its destination has no free slot at or above the cursor (even if lower slots are free),
so drops inherit the cursor's dependence on the decomposition.  The runs here use 64 slots for 12 particles per cell, which
keeps every step free of drops (asserted on the oracle side).

Two transports: (a) two "ranks" on ONE GPU exchanging planes through jp_halo_pack / jp_halo_unpack and a device copy --
runs on the single-GPU box, covers the decomposition, the halo planes, the hand-off's re-classification of rewritten
planes and the shell / interior split; (b) two processes on two GPUs through jp_halo_exchange over NCCL (sequential and
overlapped, halo.advection_with_halo) -- needs 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DX = 1.0 / 64.0
NXCELL, SLOTS, MINX, SEED = 12, 64, 6, 11     # 64 slots: no particle is dropped in these runs (see the module docstring)


def _ranges(J, i0, n):
    """StepRange (exact) staggered grids of a block of n cells starting at global cell i0 (per dimension)"""
    xv = [J.StepRange(i0[d] * DX, (i0[d] + n[d]) * DX, n[d] + 1) for d in range(3)]
    xc = [J.StepRange(i0[d] * DX + DX / 2, (i0[d] + n[d]) * DX - DX / 2, n[d]) for d in range(3)]
    xg = [J.expand_range(c) for c in xc]
    return tuple(tuple(xv[d] if d == comp else xg[d] for d in range(3)) for comp in range(3))


def _velocity(grid_vel):
    V = []
    for comp in range(3):
        x = np.asarray(grid_vel[comp][0])[None, None, :]
        z = np.asarray(grid_vel[comp][2])[:, None, None]
        shape = tuple(len(grid_vel[comp][d]) for d in (2, 1, 0))
        v = 250.0 * np.sin(np.pi * x) * np.cos(np.pi * z) if comp == 0 else (-250.0 * np.cos(np.pi * x) * np.sin(np.pi * z) if comp == 2 else np.zeros((1, 1, 1)))
        V.append(np.ascontiguousarray(np.broadcast_to(v, shape), dtype=np.float64))
    return V


class Problem:
    def __init__(self, dims, nloc):
        import justpic.jl_b200 as J
        from oracle.oracle import Oracle
        self.J, self.dims, self.nloc = J, dims, nloc
        self.nglob = tuple(dims[d] * (nloc[d] - 2) + 2 if dims[d] > 1 else nloc[d] for d in range(3))
        assert all(dims[d] == 1 or (nloc[d] - 2) % 3 == 0 for d in range(3)), "owned width must be a multiple of 3 (sweep colours)"
        self.gv = _ranges(J, (0, 0, 0), self.nglob)
        xi_vel = tuple(tuple(np.asarray(x, dtype=np.float64) for x in g) for g in self.gv)
        xvi = tuple(xi_vel[i][i] for i in range(3))
        xci = (xi_vel[1][0][1:-1].copy(), xi_vel[0][1][1:-1].copy(), xi_vel[0][2][1:-1].copy())
        self.o = Oracle(xvi, xci, xi_vel, SLOTS, True)
        self.co, self.idx = self.o.init_particles(NXCELL, SEED)
        self.V = _velocity(self.gv)
        self.dt = 0.7 * DX / 250.0
        self.fields = [np.where(self.idx > 0, 3.0 * self.co[0] + self.co[2], 0.0), np.where(self.idx > 0, 1.0 + (self.co[0] < self.co[2]), 0.0)]
        zv = xvi[2]
        self.T = np.ascontiguousarray(np.broadcast_to(zv[:, None, None], tuple(n + 1 for n in reversed(self.nglob))))

    def offset(self, coords):
        return tuple(coords[d] * (self.nloc[d] - 2) if self.dims[d] > 1 else 0 for d in range(3))

    def block(self, a, i0, plus=0):
        sl = tuple(slice(i0[d], i0[d] + self.nloc[d] + plus) for d in (2, 1, 0))
        return np.ascontiguousarray(a[(slice(None),) + sl] if a.ndim == 4 else a[sl])

    def make_rank(self, topo, device):
        """Particles of one rank, initialised with its block of the global state"""
        J = self.J
        i0 = self.offset(topo.coords())
        gv = _ranges(J, i0, self.nloc)
        p = J.init_particles(J.CUDABackend, NXCELL, SLOTS, MINX, *gv, seed=SEED, device=device)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        for d in range(3):
            p.coords[d].copy_(dev(self.block(self.co[d], i0)))
            # the block's grid vectors are the global doubles
            assert np.array_equal(np.asarray(gv[d][d]), np.asarray(self.gv[d][d])[i0[d]:i0[d] + self.nloc[d] + 1])
        p.index.copy_(dev(self.block(self.idx, i0)))
        args = tuple(dev(self.block(f, i0)) for f in self.fields)
        Vd = [dev(v) for v in _velocity(gv)]
        for c in range(3):                              # ... and its velocity arrays are slices of the global ones
            ext = tuple(len(gv[c][d]) for d in (2, 1, 0))
            sl = tuple(slice(i0[d], i0[d] + ext[k]) for k, d in enumerate((2, 1, 0)))
            assert np.array_equal(Vd[c].cpu().numpy(), self.V[c][sl])
        return p, args, Vd, i0

    def oracle_step(self):
        self.o.advect(self.co, self.idx, 1, 0.5, self.V, self.dt)
        st = self.o.move(self.co, self.idx, self.fields)
        assert st[1] == 0, "no drops expected (they depend on the slot cursor)"
        oT = np.empty_like(self.T)
        self.o.particle2grid(self.co, self.idx, oT, self.fields[0])
        return oT

    def owned(self, topo):
        """per dimension: local cell range [lo, hi) this rank owns"""
        out = []
        for d in range(3):
            lo = 1 if topo.neighbor(d, -1) is not None else 0
            hi = self.nloc[d] - 1 if topo.neighbor(d, +1) is not None else self.nloc[d]
            out.append((lo, hi))
        return out

    def compare(self, topo, i0, state, F, oT, what):
        """state: host arrays [x, y, z, fields..., index] of one rank; F: its particle2grid result"""
        own = self.owned(topo)
        lsl = (slice(None),) + tuple(slice(own[d][0], own[d][1]) for d in (2, 1, 0))
        gsl = (slice(None),) + tuple(slice(i0[d] + own[d][0], i0[d] + own[d][1]) for d in (2, 1, 0))
        gstate = self.co + self.fields + [self.idx]
        lx, gx = state[0][lsl], gstate[0][gsl]
        lo_, go_ = np.argsort(lx, axis=0, kind="stable"), np.argsort(gx, axis=0, kind="stable")      # NaN (dead slots) sort last
        assert np.array_equal(state[-1][lsl].sum(axis=0), gstate[-1][gsl].sum(axis=0)), f"{what}: live counts per owned cell differ"
        dead = np.isnan(np.take_along_axis(lx, lo_, axis=0))                 # dead slots carry no particle: content not compared
        for k in range(len(state) - 1):
            a = np.take_along_axis(state[k][lsl], lo_, axis=0); b = np.take_along_axis(gstate[k][gsl], go_, axis=0)
            a = np.where(dead, 0.0, a); b = np.where(dead, 0.0, b)
            assert np.array_equal(a, b, equal_nan=True), f"{what}: array {k} differs from the global run in {int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())} entries"
        slot_exact = float(np.mean([np.array_equal(state[k][lsl], gstate[k][gsl], equal_nan=True) for k in range(len(state))]))
        # particle2grid! at the nodes that touch no halo cell (the reference strips T[2:end-1] before gathering, :91)
        nsl, gnsl = [], []
        for d in (2, 1, 0):
            lo = 2 if topo.neighbor(d, -1) is not None else 0
            hi = self.nloc[d] - 1 if topo.neighbor(d, +1) is not None else self.nloc[d] + 1
            nsl.append(slice(lo, hi)); gnsl.append(slice(i0[d] + lo, i0[d] + hi))
        np.testing.assert_allclose(F[tuple(nsl)], oT[tuple(gnsl)], rtol=1e-12, atol=1e-12, err_msg=f"{what}: particle2grid")
        return slot_exact


def _host_state(p, args):
    return [c.cpu().numpy() for c in p.coords] + [a.cpu().numpy() for a in args] + [p.index.cpu().numpy()]


@pytest.mark.parametrize("dims,nloc", [((1, 1, 2), (34, 9, 32)), ((2, 1, 1), (35, 9, 10)), ((1, 2, 2), (33, 8, 11))],
                         ids=["z-split", "x-split", "yz-split"])
@pytest.mark.parametrize("split", [False, True], ids=["advect-then-exchange", "shell-exchange-interior"])
def test_decomposed_on_one_gpu_equals_global_oracle(dims, nloc, split):
    import justpic.jl_b200 as J
    from justpic.jl_b200 import halo
    pb = Problem(dims, nloc)
    world = int(np.prod(dims))
    ranks = []
    for r in range(world):
        topo = halo.CartesianTopology(dims, r)
        p, args, Vd, i0 = pb.make_rank(topo, torch.device("cuda", 0))
        ranks.append((topo, p, args, Vd, i0))

    def exchange():
        """update_cell_halo! between the co-located ranks: pack kernels + device copies instead of NCCL"""
        for dim in range(3):
            sends = {}
            for topo, p, args, _, _ in ranks:
                arrays = tuple(p.coords) + tuple(args)
                nb = halo.plane_bytes(p.ncells, p.max_xcell, dim, len(arrays))
                for side, plane in ((-1, 1), (+1, p.ncells[dim] - 2)):
                    if topo.neighbor(dim, side) is not None:
                        b = torch.empty(nb, dtype=torch.uint8, device="cuda")
                        halo._cuda_pack(p, dim, plane, arrays, b)
                        sends[(topo.rank, side)] = b
            for topo, p, args, _, _ in ranks:
                arrays = tuple(p.coords) + tuple(args)
                left, right = topo.neighbor(dim, -1), topo.neighbor(dim, +1)
                if left is not None:
                    halo._cuda_unpack(p, dim, 0, arrays, sends[(left, +1)])
                if right is not None:
                    halo._cuda_unpack(p, dim, p.ncells[dim] - 1, arrays, sends[(right, -1)])

    Fs = [torch.empty(tuple(n + 1 for n in reversed(nloc)), dtype=torch.float64, device="cuda") for _ in ranks]
    for it in range(4):
        if split:
            for topo, p, args, Vd, _ in ranks:
                J.advection(p, J.RungeKutta2(), Vd, pb.dt, classify=True, region="shell")
            exchange()
            for topo, p, args, Vd, _ in ranks:
                J.advection(p, J.RungeKutta2(), Vd, pb.dt, region="interior")
        else:
            for topo, p, args, Vd, _ in ranks:
                J.advection(p, J.RungeKutta2(), Vd, pb.dt, classify=(it % 2 == 0))
            exchange()
        for k, (topo, p, args, Vd, _) in enumerate(ranks):
            J.move_particles(p, args)
            assert J.last_move_path(p) == "plan"
            J.particle2grid(Fs[k], args[0], p)
        oT = pb.oracle_step()
        for k, (topo, p, args, Vd, i0) in enumerate(ranks):
            pb.compare(topo, i0, _host_state(p, args), Fs[k].cpu().numpy(), oT, f"step {it} rank {topo.rank}")


# ------------------------------------------------------------------------------------------ two GPUs, NCCL
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, dims, nloc, overlap, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)         # only carries the NCCL id: the exchange is jp_halo_exchange
    try:
        import justpic.jl_b200 as J
        from justpic.jl_b200 import halo
        pb = Problem(dims, nloc)
        topo = halo.CartesianTopology(dims, rank)
        comm = halo.create_comm(device=f"cuda:{rank}")
        p, args, Vd, i0 = pb.make_rank(topo, torch.device("cuda", rank))
        F = torch.empty(tuple(n + 1 for n in reversed(nloc)), dtype=torch.float64, device=f"cuda:{rank}")
        vmax = torch.tensor([float(np.abs(v.cpu().numpy()).max()) for v in Vd], device=f"cuda:{rank}", dtype=torch.float64)
        halo.allreduce_max(comm, vmax)                                    # the reference's dt reduction (:71)
        assert float(vmax.max()) <= 250.0
        snaps = []
        for it in range(4):
            if overlap:
                halo.advection_with_halo(p, J.RungeKutta2(), Vd, pb.dt, args, topo, classify=True, comm=comm)
            else:
                J.advection(p, J.RungeKutta2(), Vd, pb.dt, classify=(it % 2 == 0))
                halo.update_cell_halo(p, args, topo, comm=comm)
            J.move_particles(p, args)
            assert J.last_move_path(p) == "plan"
            J.particle2grid(F, args[0], p)
            torch.cuda.synchronize()
            snaps.append((_host_state(p, args), F.cpu().numpy()))
        torch.save(snaps, out + f".{rank}")
        comm.destroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("dims,nloc", [((1, 1, 2), (34, 9, 32)), ((2, 1, 1), (35, 9, 10))], ids=["z-split", "x-split"])
@pytest.mark.parametrize("overlap", [False, True], ids=["sequential", "overlapped"])
def test_decomposed_on_two_gpus_equals_global_oracle(tmp_path, dims, nloc, overlap):
    import torch.multiprocessing as mp
    from justpic.jl_b200 import halo
    out = str(tmp_path / "dec")
    mp.spawn(_worker, args=(2, _free_port(), dims, nloc, overlap, out), nprocs=2, join=True)
    pb = Problem(dims, nloc)
    snaps = [torch.load(out + f".{k}", weights_only=False) for k in range(2)]
    for it in range(4):
        oT = pb.oracle_step()
        for k in range(2):
            topo = halo.CartesianTopology(dims, k)
            state, F = snaps[k][it]
            pb.compare(topo, pb.offset(topo.coords()), state, F, oT, f"step {it} rank {k}")
