"""update_cell_halo! transport schedule on CPU: world_size-2 gloo processes run
justpic.jl_b200.halo.exchange_planes with torch-slicing pack/unpack (on the GPU the
same schedule drives the jp_halo_pack / jp_halo_unpack kernels over NCCL).
Checks the ImplicitGlobalGrid semantics (overlap 2, halo width 1): my plane 2 ->
left neighbour's plane n, my plane n-1 -> right neighbour's plane 1 (1-based),
dimensions processed x -> y -> z."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from justpic.jl_b200.halo import CartesianTopology, exchange_planes, plane_bytes


def test_cartesian_topology_matches_mpi_cart_layout():
    t = CartesianTopology((2, 2, 2), 5)             # row-major like MPI_Cart_create: rank = (cx*2 + cy)*2 + cz
    assert t.coords() == (1, 0, 1)
    assert t.rank_of((1, 0, 1)) == 5
    assert t.neighbor(0, -1) == 1 and t.neighbor(0, +1) is None
    assert t.neighbor(1, +1) == 7 and t.neighbor(2, -1) == 4
    # MPI_Dims_create order (ImplicitGlobalGrid's default) ...
    assert CartesianTopology.create(8, 3, 0, split_x_last=False).dims == (2, 2, 2)
    assert CartesianTopology.create(4, 3, 0, split_x_last=False).dims == (2, 2, 1)
    assert CartesianTopology.create(2, 3, 0, split_x_last=False).dims == (2, 1, 1)
    # ... and the default here: x (the contiguous axis of the CellArray layout) is never split
    assert CartesianTopology.create(8, 3, 0).dims == (1, 2, 4)
    assert CartesianTopology.create(4, 3, 0).dims == (1, 2, 2)
    assert CartesianTopology.create(2, 3, 0).dims == (1, 1, 2)
    p = CartesianTopology((2, 1), 0, periodic=(True, False))
    assert p.neighbor(0, -1) == 1 and p.neighbor(0, +1) == 1 and p.neighbor(1, +1) is None
    q = CartesianTopology((2, 1), 1, periodic=(False, True))         # one rank along a periodic dimension: its own neighbour
    assert q.neighbor(1, -1) == 1 and q.neighbor(1, +1) == 1 and q.neighbor(0, +1) is None and q.decomposed
    assert list(q.neighbor_table()) == [0, -1, 1, 1, -1, -1]
    assert not CartesianTopology((1, 1, 1), 0).decomposed


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _plane(t, dim, plane, ndim):
    # tensors are (S, [nz,] ny, nx): spatial dim d is tensor axis ndim - d
    return t.select(ndim - dim, plane)


def _worker(rank, world, port, dims, ncells, S, out, periodic=()):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ndim = len(ncells)
        topo = CartesianTopology(dims, rank, tuple(periodic))
        shape = (S, *reversed(ncells))
        g = torch.Generator().manual_seed(100 + rank)
        arrays = [torch.rand(shape, generator=g, dtype=torch.float64) + 10 * rank for _ in range(3)]
        index = (torch.rand(shape, generator=g) < 0.5).to(torch.uint8)
        before = [a.clone() for a in arrays] + [index.clone()]

        def pack(dim, plane, buf):
            parts = [_plane(a, dim, plane, ndim).contiguous().view(torch.uint8).reshape(-1) for a in arrays]
            parts.append(_plane(index, dim, plane, ndim).contiguous().reshape(-1))
            buf.copy_(torch.cat(parts))

        def unpack(dim, plane, buf):
            off = 0
            for a in arrays:
                dst = _plane(a, dim, plane, ndim)
                nb = dst.numel() * 8
                dst.copy_(buf[off:off + nb].view(torch.float64).reshape(dst.shape))
                off += nb
            dst = _plane(index, dim, plane, ndim)
            dst.copy_(buf[off:off + dst.numel()].reshape(dst.shape))

        sent = exchange_planes(topo, ncells, S, len(arrays), "cpu", pack, unpack)
        torch.save({"before": before, "after": arrays + [index], "sent": sent}, out + f".{rank}")
    finally:
        dist.destroy_process_group()


def _expected(dims, periodic, ncells, before):
    """update_halo! semantics simulated serially: dimensions in order, every rank's plane 0 <- left neighbour's plane n-2 and
    plane n-1 <- right neighbour's plane 1, both taken from the state BEFORE this dimension's exchange."""
    world, ndim = len(before), len(ncells)
    cur = [[a.clone() for a in before[r]] for r in range(world)]
    for dim in range(ndim):
        n = ncells[dim]
        nxt = [[a.clone() for a in cur[r]] for r in range(world)]
        for r in range(world):
            topo = CartesianTopology(dims, r, tuple(periodic))
            left, right = topo.neighbor(dim, -1), topo.neighbor(dim, +1)
            for a in range(len(cur[r])):
                if left is not None:
                    _plane(nxt[r][a], dim, 0, ndim).copy_(_plane(cur[left][a], dim, n - 2, ndim))
                if right is not None:
                    _plane(nxt[r][a], dim, n - 1, ndim).copy_(_plane(cur[right][a], dim, 1, ndim))
        cur = nxt
    return cur


@pytest.mark.parametrize("dims,ncells,periodic", [
    ((2, 1), (6, 5), ()), ((1, 2), (6, 5), ()), ((2, 1, 1), (5, 4, 6), ()), ((1, 1, 2), (5, 4, 6), ()),
    # periodic boundaries: two ranks along the dimension are each other's left AND right neighbour (message order decides which
    # plane lands where); one rank along a periodic dimension is its own neighbour (local wrap-around)
    ((2, 1), (6, 5), (True, False)), ((1, 2), (6, 5), (True, True)), ((2, 1, 1), (5, 4, 6), (True, True, False)),
    ((1, 1, 2), (5, 4, 6), (False, True, True)),
])
def test_exchange_planes_two_ranks(tmp_path, dims, ncells, periodic):
    S, world = 3, 2
    out = str(tmp_path / "halo")
    mp.spawn(_worker, args=(world, _free_port(), dims, ncells, S, out, periodic), nprocs=world, join=True)
    r = [torch.load(out + f".{k}") for k in range(world)]
    want = _expected(dims, periodic, ncells, [r[k]["before"] for k in range(world)])
    for k in range(world):
        for a in range(4):
            assert torch.equal(r[k]["after"][a], want[k][a]), f"rank {k} array {a}"
    if not periodic:
        ndim = len(ncells)
        dim = [i for i, d in enumerate(dims) if d == 2][0]
        n = ncells[dim]
        for a in range(4):
            # rank 1 is the right neighbour of rank 0 along `dim`
            assert torch.equal(_plane(r[0]["after"][a], dim, n - 1, ndim), _plane(r[1]["before"][a], dim, 1, ndim))
            assert torch.equal(_plane(r[1]["after"][a], dim, 0, ndim), _plane(r[0]["before"][a], dim, n - 2, ndim))
        assert r[0]["sent"] == r[1]["sent"] == plane_bytes(ncells, S, dim, 3)
