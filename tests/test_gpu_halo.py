"""jp_halo_pack / jp_halo_unpack kernels (single GPU) and the NCCL exchange (2 GPUs if present)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _particles(ndim, n, seed=3):
    import justpic.jl_b200 as J
    from tests.problems import make_grids
    gr = make_grids(n, ndim, True)
    return J, gr, J.init_particles(J.CUDABackend, 12, 16, 6, *gr.grid_vel, seed=seed)


@pytest.mark.parametrize("ndim,n", [(2, (9, 7)), (3, (6, 5, 7))])
def test_halo_pack_unpack_kernels(ndim, n):
    from justpic.jl_b200 import halo
    J, gr, p = _particles(ndim, n)
    pT, = J.init_cell_arrays(p, 1)
    pT.copy_(torch.rand_like(pT))
    arrays = tuple(p.coords) + (pT,)
    for dim in range(ndim):
        nb = halo.plane_bytes(p.ncells, p.max_xcell, dim, len(arrays))
        buf = torch.empty(nb, dtype=torch.uint8, device="cuda")
        src_plane, dst_plane = 1, p.ncells[dim] - 1
        halo._cuda_pack(p, dim, src_plane, arrays, buf)
        # reference packing with torch slicing: arrays in order (slot-major, plane cells in memory order), then index bytes
        parts = [a.select(ndim - dim, src_plane).contiguous().view(torch.uint8).reshape(-1) for a in arrays]
        parts.append(p.index.select(ndim - dim, src_plane).contiguous().reshape(-1))
        assert torch.equal(buf, torch.cat(parts))
        before = [a.clone() for a in arrays] + [p.index.clone()]
        halo._cuda_unpack(p, dim, dst_plane, arrays, buf)
        for a, b in zip(list(arrays) + [p.index], before):
            got = a.select(ndim - dim, dst_plane)
            want = b.select(ndim - dim, src_plane)
            assert torch.equal(torch.nan_to_num(got.double(), nan=-1.0), torch.nan_to_num(want.double(), nan=-1.0))
            keep = [i for i in range(p.ncells[dim]) if i != dst_plane]
            idx = torch.tensor(keep, device="cuda")
            assert torch.equal(torch.nan_to_num(a.index_select(ndim - dim, idx).double(), nan=-1.0),
                               torch.nan_to_num(b.index_select(ndim - dim, idx).double(), nan=-1.0))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _nccl_worker(rank, world, port, out, use_comm=True):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import justpic.jl_b200 as J
        from justpic.jl_b200.halo import CartesianTopology, update_cell_halo
        from tests.problems import make_grids
        gr = make_grids((6, 5, 7), 3, True)
        p = J.init_particles(J.CUDABackend, 12, 16, 6, *gr.grid_vel, seed=10 + rank, device=f"cuda:{rank}")
        pT, = J.init_cell_arrays(p, 1)
        pT.fill_(float(rank) + 0.5)
        before = [c.clone().cpu() for c in p.coords] + [pT.clone().cpu(), p.index.clone().cpu()]
        topo = CartesianTopology((2, 1, 1), rank)
        from justpic.jl_b200.halo import create_comm
        comm = create_comm(device=f"cuda:{rank}") if use_comm else None      # jp_halo_exchange (C, NCCL) / torch.distributed transport
        sent = update_cell_halo(p, (pT,), topo, comm=comm)
        torch.cuda.synchronize()
        if comm is not None:
            comm.destroy()
        after = [c.cpu() for c in p.coords] + [pT.cpu(), p.index.cpu()]
        torch.save({"before": before, "after": after, "sent": sent}, out + f".{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("use_comm", [True, False], ids=["jp_halo_exchange", "torch.distributed"])
def test_update_cell_halo_nccl_two_gpus(tmp_path, use_comm):
    import torch.multiprocessing as mp
    out = str(tmp_path / "halo")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out, use_comm), nprocs=2, join=True)
    r = [torch.load(out + f".{k}") for k in range(2)]
    nx = 6
    eq = lambda a, b: torch.equal(torch.nan_to_num(a.double(), nan=-1.0), torch.nan_to_num(b.double(), nan=-1.0))
    for a in range(5):
        assert eq(r[0]["after"][a][..., nx - 1], r[1]["before"][a][..., 1])
        assert eq(r[1]["after"][a][..., 0], r[0]["before"][a][..., nx - 2])
        assert eq(r[0]["after"][a][..., : nx - 1], r[0]["before"][a][..., : nx - 1])
        assert eq(r[1]["after"][a][..., 1:], r[1]["before"][a][..., 1:])


# ---------------------------------------------------------------------------------------------------------------
# advection! split into shell + interior launches (jp_advect_region), the basis of the halo overlap
@pytest.mark.parametrize("classify", [False, True])
@pytest.mark.parametrize("g", [(2, (70, 19), True), (3, (40, 9, 7), True), (3, (33, 6, 5), False), (3, (70, 11, 6), True)],
                         ids=lambda g: f"{g[0]}D-{g[1]}-{'range' if g[2] else 'vector'}")
def test_advection_shell_plus_interior_equals_advection(g, classify):
    """The two region launches together are one advection!: bit-identical coordinates for every integrator, and the
    advection -> move hand-off they leave drives move_particles! to the oracle's result."""
    import numpy as np
    import justpic.jl_b200 as J
    from tests.problems import cfl_dt, stream_velocity
    from tests.test_gpu_parity import Twin, dev
    ndim, n, uniform = g
    t = Twin(ndim, n, uniform=uniform)
    V = stream_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(t.gr, V, 0.8)
    pT, = J.init_cell_arrays(t.p, 1)
    opT = np.zeros_like(t.co[0])
    methods = [(J.RungeKutta2(), 1, 0.5), (J.RungeKutta4(), 2, 0.0), (J.Euler(), 0, 0.0), (J.RungeKutta2(2 / 3), 1, 2 / 3)]
    for it, m in enumerate(methods):
        J.advection(t.p, m[0], Vd, dt, classify=classify, region="shell")
        J.advection(t.p, m[0], Vd, dt, region="interior")
        t.o.advect(t.co, t.idx, m[1], m[2], V, dt)
        t.check_state(f"step {it} advection (shell + interior)")
        J.move_particles(t.p, (pT,)); st = t.o.move(t.co, t.idx, [opT])
        t.check_state(f"step {it} move_particles", (pT,), (opT,))
        assert J.move_stats(t.p) == st
        if classify:
            assert J.last_move_classify(t.p) == "handoff"
    with pytest.raises(ValueError):                    # interior without shell (JP_ERR_INVALID)
        J.advection(t.p, methods[0][0], Vd, dt, region="interior")
    J.advection(t.p, methods[0][0], Vd, dt, region="shell")
    with pytest.raises(ValueError):                    # shell not completed
        J.advection(t.p, methods[0][0], Vd, dt)
    J.advection(t.p, methods[0][0], Vd, dt)            # the failed call reset the split: a plain advection works again


def _overlap_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import numpy as np
        import justpic.jl_b200 as J
        from justpic.jl_b200.halo import CartesianTopology, advection_with_halo, update_cell_halo, create_comm
        from bench import stream_velocity_np
        comm = create_comm(device=f"cuda:{rank}")
        # a real block decomposition (2 x 1 x 1, overlap 2 as ImplicitGlobalGrid): the particles a rank receives in its halo
        # cells lie in those cells, so move_particles! re-buckets over <= 1 cell (longer moves are racy in the reference itself)
        topo = CartesianTopology((2, 1, 1), rank)
        n = (72, 14, 10)                                 # bricks are 32 x 4 x 2 cells: x [32,64), y [4,12), z [2,8) are interior
        xv, xc = [], []
        for d in range(3):
            nglob = topo.dims[d] * (n[d] - 2) + 2 if topo.dims[d] > 1 else n[d]
            dx = 1.0 / nglob
            i0 = topo.coords()[d] * (n[d] - 2) if topo.dims[d] > 1 else 0
            xv.append(J.LinRange(i0 * dx, (i0 + n[d]) * dx, n[d] + 1))
            xc.append(J.LinRange(i0 * dx + dx / 2, (i0 + n[d]) * dx - dx / 2, n[d]))
        xg = [J.expand_range(c) for c in xc]
        grid_vel = tuple(tuple(xv[d] if d == comp else xg[d] for d in range(3)) for comp in range(3))
        Vd = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in stream_velocity_np(grid_vel)]
        dt = 0.7 * min(float(xv[d][1] - xv[d][0]) for d in range(3)) / 250.0
        res = []
        for overlapped in (False, True):
            p = J.init_particles(J.CUDABackend, 12, 24, 6, *grid_vel, seed=10 + rank, device=f"cuda:{rank}")
            pT, = J.init_cell_arrays(p, 1)
            pT.copy_(p.coords[0] * 3.0)
            bufs = {}
            for it in range(4):
                if overlapped:
                    advection_with_halo(p, J.RungeKutta2(), Vd, dt, (pT,), topo, buffers=bufs, classify=True, comm=comm)
                else:
                    J.advection(p, J.RungeKutta2(), Vd, dt, classify=True)
                    update_cell_halo(p, (pT,), topo, buffers=bufs, comm=comm)
                J.move_particles(p, (pT,))
                assert J.last_move_classify(p) == "handoff"
            torch.cuda.synchronize()
            res.append([c.cpu() for c in p.coords] + [pT.cpu(), p.index.cpu()])
        torch.save(res, out + f".{rank}")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_advection_with_halo_overlap_two_gpus(tmp_path):
    """advection_with_halo (shell launch, exchange on a side stream behind the interior launch) == advection! then
    update_cell_halo!, bit for bit, over coupled steps with move_particles! and the hand-off."""
    import torch.multiprocessing as mp
    out = str(tmp_path / "ovl")
    mp.spawn(_overlap_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for k in range(2):
        seq, ovl = torch.load(out + f".{k}")
        for a, b in zip(seq, ovl):
            assert torch.equal(torch.nan_to_num(a.double(), nan=-1.0), torch.nan_to_num(b.double(), nan=-1.0))


# ---------------------------------------------------------------------------------------------------------------
# jp_halo_exchange / jp_halo_exchange_grid on ONE GPU: a rank that is its own neighbour (periodic, one rank along the dimension)
# wraps around locally -- the library's pack kernels, plane numbers and x -> y -> z order without a second GPU
@pytest.mark.parametrize("ndim,n,periodic", [(2, (9, 7), (True, True)), (3, (6, 5, 7), (True, False, True)), (3, (8, 6, 5), (False, True, False))])
def test_halo_exchange_periodic_self_wrap(ndim, n, periodic):
    from justpic.jl_b200.halo import CartesianTopology, update_cell_halo, update_halo
    J, gr, p = _particles(ndim, n)
    pT, = J.init_cell_arrays(p, 1)
    pT.copy_(torch.rand_like(pT))
    topo = CartesianTopology((1,) * ndim, 0, periodic)
    arrays = list(p.coords) + [pT, p.index]
    want = [a.clone() for a in arrays]
    for dim in range(ndim):                       # update_halo! semantics, dimensions in sequence
        if not periodic[dim]:
            continue
        ax, nn = ndim - dim, n[dim]
        for w in want:
            src_lo, src_hi = w.select(ax, nn - 2).clone(), w.select(ax, 1).clone()
            w.select(ax, 0).copy_(src_lo)
            w.select(ax, nn - 1).copy_(src_hi)
    update_cell_halo(p, (pT,), topo)
    for a, w in zip(arrays, want):
        assert torch.equal(torch.nan_to_num(a.double(), nan=-1.0), torch.nan_to_num(w.double(), nan=-1.0))
    # a staggered grid array (overlap 2 + extent - cells): planes ol-1 / ext-ol -> planes ext-1 / 0 (0-based)
    for plus in ((1, 2, 2), (2, 1, 2), (1, 1, 1), (0, 0, 0)):
        ext = tuple(n[d] + plus[d] for d in range(ndim))
        A = torch.rand(tuple(reversed(ext)), dtype=torch.float64, device="cuda")
        W = A.clone()
        for dim in range(ndim):
            if not periodic[dim]:
                continue
            ax, ol = ndim - 1 - dim, 2 + plus[dim]
            lo, hi = W.select(ax, ext[dim] - ol).clone(), W.select(ax, ol - 1).clone()
            W.select(ax, 0).copy_(lo)
            W.select(ax, ext[dim] - 1).copy_(hi)
        update_halo(p, A, topo)
        assert torch.equal(A, W)
