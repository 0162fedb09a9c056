"""update_cell_halo! for block-decomposed domains (one process per GPU).

Reference semantics (src/CellArrays/ImplicitGlobalGrid.jl:36-41, i.e.
ImplicitGlobalGrid.update_halo! with overlap 2 / halo width 1, applied to every
CellArray): per dimension, in the order x -> y -> z,

    my cell-plane 2      -> left  neighbour's plane n      (1-based)
    my cell-plane n - 1  -> right neighbour's plane 1

so edges/corners propagate through the sequential dimensions.  Particle
migration is implicit: the neighbour's boundary cells land in my halo cells and
``move_particles`` re-buckets whatever now lies in my interior
(scripts/temperature_advection3D_MPI.jl:83-91).

Here the planes of *all* listed CellArrays (coords, particle fields, index) are
gathered by one CUDA kernel into one contiguous buffer per face
(``jp_halo_pack``), exchanged with NCCL point-to-point send/recv through
``torch.distributed`` and scattered by ``jp_halo_unpack`` -- one message per
face instead of one per array.  Rank layout follows MPI_Cart_create
(row-major: the last dimension varies fastest), as ImplicitGlobalGrid's does.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _cabi
from .api import Particles, _ptr_array, _stream, advection

__all__ = ["CartesianTopology", "update_cell_halo", "exchange_planes", "advection_with_halo", "join_halo"]


@dataclass(frozen=True)
class CartesianTopology:
    dims: Tuple[int, ...]
    rank: int
    periodic: Tuple[bool, ...] = ()

    def __post_init__(self):
        if not self.periodic:
            object.__setattr__(self, "periodic", tuple(False for _ in self.dims))
        size = 1
        for d in self.dims:
            size *= d
        if not (0 <= self.rank < size):
            raise ValueError("rank outside the Cartesian topology")

    @property
    def size(self) -> int:
        s = 1
        for d in self.dims:
            s *= d
        return s

    def coords(self, rank: Optional[int] = None) -> Tuple[int, ...]:
        r = self.rank if rank is None else rank
        out = []
        for d in reversed(self.dims):
            out.append(r % d)
            r //= d
        return tuple(reversed(out))

    def rank_of(self, coords: Sequence[int]) -> int:
        r = 0
        for c, d in zip(coords, self.dims):
            r = r * d + c
        return r

    def neighbor(self, dim: int, side: int) -> Optional[int]:
        """side = -1 (left) / +1 (right); None at a non-periodic boundary."""
        c = list(self.coords())
        c[dim] += side
        if not (0 <= c[dim] < self.dims[dim]):
            if not self.periodic[dim] or self.dims[dim] == 1:
                return None
            c[dim] %= self.dims[dim]
        return self.rank_of(c)

    @staticmethod
    def create(world_size: int, ndim: int, rank: int) -> "CartesianTopology":
        """MPI_Dims_create-like balanced factorisation (largest factors first)."""
        dims = [1] * ndim
        n = world_size
        f = 2
        factors = []
        while n > 1:
            while n % f == 0:
                factors.append(f)
                n //= f
            f += 1
        for p in sorted(factors, reverse=True):
            i = dims.index(min(dims))
            dims[i] *= p
        dims.sort(reverse=True)
        return CartesianTopology(tuple(dims), rank)


def _cuda_pack(particles: Particles, dim: int, plane: int, arrays, buf: torch.Tensor) -> None:
    _cabi.check(_cabi.load().jp_halo_pack(C.c_void_p(particles._ctx), dim, plane, _ptr_array(arrays), len(arrays),
                                          C.c_void_p(particles.index.data_ptr()), C.c_void_p(buf.data_ptr()), _stream()),
                "jp_halo_pack")


def _cuda_unpack(particles: Particles, dim: int, plane: int, arrays, buf: torch.Tensor) -> None:
    _cabi.check(_cabi.load().jp_halo_unpack(C.c_void_p(particles._ctx), dim, plane, _ptr_array(arrays), len(arrays),
                                            C.c_void_p(particles.index.data_ptr()), C.c_void_p(buf.data_ptr()), _stream()),
                "jp_halo_unpack")


def plane_bytes(ncells: Sequence[int], S: int, dim: int, narrays: int) -> int:
    m = 1
    for d, n in enumerate(ncells):
        if d != dim:
            m *= n
    return m * S * (8 * narrays + 1)


def exchange_planes(topo: CartesianTopology, ncells: Sequence[int], S: int, narrays: int, device,
                    pack: Callable[[int, int, torch.Tensor], None], unpack: Callable[[int, int, torch.Tensor], None],
                    group=None, buffers: Optional[dict] = None) -> int:
    """The transport schedule, independent of how planes are packed: for each
    dimension pack planes 1 and n-2 (0-based), send them left / right, receive
    into planes n-1 / 0, unpack.  ``pack(dim, plane, buf)`` / ``unpack`` are the
    CUDA kernels in production and torch slicing in the CPU (gloo) tests.
    Returns the number of bytes this rank sent."""
    sent = 0
    ndim = len(ncells)
    for dim in range(ndim):
        left, right = topo.neighbor(dim, -1), topo.neighbor(dim, +1)
        if left is None and right is None:
            continue
        nb = plane_bytes(ncells, S, dim, narrays)
        n = ncells[dim]

        def buf(tag):
            key = (dim, tag)
            if buffers is not None and key in buffers:
                return buffers[key]
            b = torch.empty(nb, dtype=torch.uint8, device=device)
            if buffers is not None:
                buffers[key] = b
            return b

        ops, recvs = [], []
        if left is not None:
            sl, rl = buf("send_l"), buf("recv_l")
            pack(dim, 1, sl)
            ops.append(dist.P2POp(dist.isend, sl, left, group=group))
            ops.append(dist.P2POp(dist.irecv, rl, left, group=group))
            recvs.append((0, rl))
            sent += nb
        if right is not None:
            sr, rr = buf("send_r"), buf("recv_r")
            pack(dim, n - 2, sr)
            ops.append(dist.P2POp(dist.isend, sr, right, group=group))
            ops.append(dist.P2POp(dist.irecv, rr, right, group=group))
            recvs.append((n - 1, rr))
            sent += nb
        if left is not None and left == right and left == topo.rank:
            # periodic with a single rank along this dim: local copy
            unpack(dim, n - 1, sl)
            unpack(dim, 0, sr)
            continue
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for plane, b in recvs:
            unpack(dim, plane, b)
    return sent


def update_cell_halo(particles: Particles, args=(), topo: Optional[CartesianTopology] = None, group=None,
                     buffers: Optional[dict] = None) -> int:
    """``update_cell_halo!(particles.coords..., args..., particles.index)``
    (src/CellArrays/ImplicitGlobalGrid.jl:36-41): refresh the 1-cell halo ring of the
    particle coordinates, the listed particle fields and the occupancy mask."""
    if topo is None or topo.size == 1 and not any(topo.periodic):
        return 0
    p = particles
    arrays = tuple(p.coords) + tuple(args)
    with torch.cuda.device(p.device):
        return exchange_planes(
            topo, p.ncells, p.max_xcell, len(arrays), p.device,
            lambda dim, plane, b: _cuda_pack(p, dim, plane, arrays, b),
            lambda dim, plane, b: _cuda_unpack(p, dim, plane, arrays, b),
            group=group, buffers=buffers)


_SIDE_STREAMS: dict = {}
_PENDING: dict = {}          # id(particles) -> event recorded on the side stream after the last unpack


def join_halo(particles: Particles) -> None:
    """Make the current stream wait for an exchange started by ``advection_with_halo(..., join=False)``."""
    ev = _PENDING.pop(id(particles), None)
    if ev is not None:
        with torch.cuda.device(particles.device):
            torch.cuda.current_stream().wait_event(ev)


def advection_with_halo(particles: Particles, method, V, dt: float, args=(), topo: Optional[CartesianTopology] = None, group=None,
                        buffers: Optional[dict] = None, classify: Optional[bool] = None, join: bool = True) -> int:
    """``advection!(particles, method, V, dt)`` followed by ``update_cell_halo!(coords..., args..., index)``
    (scripts/temperature_advection3D_MPI.jl:83-91) with the exchange hidden behind the advection: the bricks holding the two
    outermost cell layers are advected first (``jp_advect_region`` SHELL), the planes are then packed, sent over NCCL and unpacked
    on a high-priority side stream while the current stream advects the interior bricks; the current stream waits for the side
    stream before returning control to the next call (``move_particles``).  Same results as the two calls in sequence."""
    if topo is None or topo.size == 1 and not any(topo.periodic):
        advection(particles, method, V, dt, classify=classify)
        return 0
    p = particles
    with torch.cuda.device(p.device):
        main = torch.cuda.current_stream()
        side = _SIDE_STREAMS.get(p.device)
        if side is None:
            side = _SIDE_STREAMS[p.device] = torch.cuda.Stream(device=p.device, priority=-1)
        advection(p, method, V, dt, classify=classify, region="shell")
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            sent = update_cell_halo(p, args, topo, group=group, buffers=buffers)
            done = torch.cuda.Event()
            done.record(side)
        advection(p, method, V, dt, region="interior")
        if join:
            main.wait_event(done)
        else:
            _PENDING[id(p)] = done          # the caller joins with join_halo(p) (bench.py: to time what is left of the exchange)
    return sent
