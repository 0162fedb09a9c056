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
    assert CartesianTopology.create(8, 3, 0).dims == (2, 2, 2)
    assert CartesianTopology.create(4, 3, 0).dims == (2, 2, 1)
    assert CartesianTopology.create(2, 3, 0).dims == (2, 1, 1)
    p = CartesianTopology((2, 1), 0, periodic=(True, False))
    assert p.neighbor(0, -1) == 1 and p.neighbor(0, +1) == 1 and p.neighbor(1, +1) is None


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _plane(t, dim, plane, ndim):
    # tensors are (S, [nz,] ny, nx): spatial dim d is tensor axis ndim - d
    return t.select(ndim - dim, plane)


def _worker(rank, world, port, dims, ncells, S, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ndim = len(ncells)
        topo = CartesianTopology(dims, rank)
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


@pytest.mark.parametrize("dims,ncells", [((2, 1), (6, 5)), ((1, 2), (6, 5)), ((2, 1, 1), (5, 4, 6)), ((1, 1, 2), (5, 4, 6))])
def test_exchange_planes_two_ranks(tmp_path, dims, ncells):
    S, world = 3, 2
    out = str(tmp_path / "halo")
    mp.spawn(_worker, args=(world, _free_port(), dims, ncells, S, out), nprocs=world, join=True)
    r = [torch.load(out + f".{k}") for k in range(world)]
    ndim = len(ncells)
    dim = [i for i, d in enumerate(dims) if d == 2][0]
    n = ncells[dim]
    for a in range(4):
        # rank 1 is the right neighbour of rank 0 along `dim`
        # rank0.plane[n-1] <- rank1.plane[1] (before);  rank1.plane[0] <- rank0.plane[n-2] (before)
        assert torch.equal(_plane(r[0]["after"][a], dim, n - 1, ndim), _plane(r[1]["before"][a], dim, 1, ndim))
        assert torch.equal(_plane(r[1]["after"][a], dim, 0, ndim), _plane(r[0]["before"][a], dim, n - 2, ndim))
        # everything else untouched
        for k, untouched in ((0, list(range(n - 1))), (1, list(range(1, n)))):
            for pl in untouched:
                assert torch.equal(_plane(r[k]["after"][a], dim, pl, ndim), _plane(r[k]["before"][a], dim, pl, ndim))
    assert r[0]["sent"] == r[1]["sent"] == plane_bytes(ncells, S, dim, 3)
