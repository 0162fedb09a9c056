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
gathered by one CUDA kernel into one contiguous buffer per face, exchanged with
NCCL point-to-point send/recv and scattered back -- one message per face instead
of one per array.  The whole exchange (pack kernels, the x -> y -> z schedule,
``ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd``) is ONE library call,
``jp_halo_exchange`` (csrc/jp_halo_nccl.cuh): this module only builds the
communicator (``create_comm``: the 128-byte NCCL id travels over whatever
``torch.distributed`` group exists) and the neighbour table.  ``exchange_planes``
below is the same schedule over ``torch.distributed`` send/recv with pluggable
pack / unpack: it is what the CPU (gloo) tests run, and the transport of last
resort when no communicator is given.  Rank layout follows MPI_Cart_create
(row-major: the last dimension varies fastest), as ImplicitGlobalGrid's does; the
default factorisation (``CartesianTopology.create``) never splits x before y and z
are split: an x face is one 8-byte element per 32-byte sector and costs 5x a y / z face.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _cabi
from .api import Particles, _ptr_array, _stream, advection

__all__ = ["CartesianTopology", "Comm", "create_comm", "update_cell_halo", "update_halo", "allreduce_max", "exchange_planes",
           "advection_with_halo", "join_halo"]


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
            if not self.periodic[dim]:
                return None
            c[dim] %= self.dims[dim]          # one rank along a periodic dimension: the rank is its own neighbour
        return self.rank_of(c)

    def neighbor_table(self):
        """(c_int32 * 6): rank of the left / right neighbour per dimension, -1 for none (the ``nbr`` of ``jp_halo_exchange``)."""
        t = (C.c_int32 * 6)(*([-1] * 6))
        for d in range(len(self.dims)):
            for k, side in enumerate((-1, +1)):
                r = self.neighbor(d, side)
                t[2 * d + k] = -1 if r is None else r
        return t

    @property
    def decomposed(self) -> bool:
        """any dimension with a neighbour (another rank, or the rank itself across a periodic boundary)"""
        return any(self.neighbor(d, s) is not None for d in range(len(self.dims)) for s in (-1, +1))

    @staticmethod
    def create(world_size: int, ndim: int, rank: int, split_x_last: bool = True) -> "CartesianTopology":
        """Balanced factorisation of the ranks over the dimensions.  ``split_x_last`` (default): the largest factors go to the
        LAST dimensions -- 2 -> (1, 1, 2), 4 -> (1, 2, 2), 8 -> (1, 2, 4) -- because a cell-plane normal to x is strided by nx in
        the CellArray layout (one useful element per 32-byte sector, and every element in a DRAM page of its own), while y / z
        planes are contiguous runs.  ``False``: MPI_Dims_create's order (largest first: 8 -> (2, 2, 2)), ImplicitGlobalGrid's default."""
        n, f, factors = world_size, 2, []
        while n > 1:
            while n % f == 0:
                factors.append(f)
                n //= f
            f += 1
        dims = [1] * ndim
        if split_x_last and ndim > 1:
            for p in sorted(factors, reverse=True):              # largest factors first, always onto the smallest of y, z
                i = min(range(1, ndim), key=lambda j: (dims[j], j))
                dims[i] *= p
            return CartesianTopology(tuple([1] + sorted(dims[1:])), rank)
        for p in sorted(factors, reverse=True):
            dims[dims.index(min(dims))] *= p
        return CartesianTopology(tuple(sorted(dims, reverse=True)), rank)


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

        sl = rl = sr = rr = None
        if left is not None:
            sl, rl = buf("send_l"), buf("recv_l")
            pack(dim, 1, sl)
            sent += nb
        if right is not None:
            sr, rr = buf("send_r"), buf("recv_r")
            pack(dim, n - 2, sr)
            sent += nb
        if left == topo.rank or right == topo.rank:
            # periodic with a single rank along this dimension: the rank is its own neighbour, wrap around locally
            unpack(dim, n - 1, sl)
            unpack(dim, 0, sr)
            continue
        # Posting order matters when left == right (periodic, two ranks along the dimension): several messages between one
        # pair of ranks match in posting order, so sends go (to-left, to-right) and receives (from-right, from-left) --
        # what I send to my left is what the peer receives from its right.
        ops = []
        if left is not None:
            ops.append(dist.P2POp(dist.isend, sl, left, group=group))
        if right is not None:
            ops.append(dist.P2POp(dist.isend, sr, right, group=group))
        if right is not None:
            ops.append(dist.P2POp(dist.irecv, rr, right, group=group))
        if left is not None:
            ops.append(dist.P2POp(dist.irecv, rl, left, group=group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if left is not None:
            unpack(dim, 0, rl)
        if right is not None:
            unpack(dim, n - 1, rr)
    return sent


class Comm:
    """An NCCL communicator owned by libjustpic_sm100a.so (``jp_comm_init``); ``handle`` is the raw ``ncclComm_t``."""

    def __init__(self, handle: int, rank: int, size: int):
        self.handle, self.rank, self.size = handle, rank, size

    def destroy(self) -> None:
        if self.handle:
            _cabi.check(_cabi.load().jp_comm_destroy(C.c_void_p(self.handle)), "jp_comm_destroy")
            self.handle = 0


def create_comm(device=None, group=None) -> Comm:
    """Collective over the ``torch.distributed`` group (any backend: it only carries the 128-byte NCCL id):
    rank 0 draws the id (``jp_comm_unique_id``), everybody joins (``jp_comm_init``)."""
    lib = _cabi.load()
    rank, size = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    idbuf = (C.c_char * 128)()
    if rank == 0:
        _cabi.check(lib.jp_comm_unique_id(idbuf), "jp_comm_unique_id")
    box = [bytes(idbuf.raw) if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    idbuf.raw = box[0]
    out = C.c_void_p()
    _cabi.check(lib.jp_comm_init(idbuf, size, rank, dev.index if dev.index is not None else torch.cuda.current_device(), C.byref(out)),
                "jp_comm_init")
    return Comm(out.value, rank, size)


def allreduce_max(comm: Comm, t: torch.Tensor) -> torch.Tensor:
    """In-place max over the ranks of a float64 device tensor (the reference scripts' ``MPI.Allreduce(..., MPI.MAX)`` for dt)."""
    if not (isinstance(t, torch.Tensor) and t.dtype == torch.float64 and t.is_cuda and t.is_contiguous()):
        raise ValueError("allreduce_max: expected a contiguous float64 device tensor")
    with torch.cuda.device(t.device):
        _cabi.check(_cabi.load().jp_allreduce_max(C.c_void_p(comm.handle), C.c_void_p(t.data_ptr()), t.numel(), _stream()), "allreduce_max")
    return t


def update_halo(particles: Particles, A: torch.Tensor, topo: CartesianTopology, comm: Optional[Comm] = None) -> None:
    """``update_halo!(A)`` (ImplicitGlobalGrid) for a plain grid array on the particles' decomposition -- a staggered
    velocity component, a vertex field: the velocity-ghost-layer exchange when V comes from a solver."""
    p = particles
    if not topo.decomposed:
        return
    if not (isinstance(A, torch.Tensor) and A.dtype == torch.float64 and A.is_cuda and A.is_contiguous() and A.dim() == p.ndim):
        raise ValueError("update_halo: expected a contiguous float64 device array with one axis per dimension")
    ext = (C.c_int32 * 3)(*(list(reversed(A.shape)) + [1] * (3 - A.dim())))
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_halo_exchange_grid(C.c_void_p(p._ctx), C.c_void_p(comm.handle if comm else None), topo.neighbor_table(),
                                                       C.c_void_p(A.data_ptr()), ext, _stream()), "update_halo")


def update_cell_halo(particles: Particles, args=(), topo: Optional[CartesianTopology] = None, group=None,
                     buffers: Optional[dict] = None, comm: Optional[Comm] = None) -> int:
    """``update_cell_halo!(particles.coords..., args..., particles.index)``
    (src/CellArrays/ImplicitGlobalGrid.jl:36-41): refresh the 1-cell halo ring of the
    particle coordinates, the listed particle fields and the occupancy mask.
    With ``comm`` (``create_comm``) -- or when every neighbour is the rank itself (periodic, undecomposed) -- this is one
    call of ``jp_halo_exchange``; without, the same schedule over ``torch.distributed`` point-to-point operations."""
    if topo is None or not topo.decomposed:
        return 0
    p = particles
    arrays = tuple(p.coords) + tuple(args)
    with torch.cuda.device(p.device):
        nbr = topo.neighbor_table()
        only_self = all(nbr[i] in (-1, topo.rank) for i in range(6))
        if comm is not None or only_self:
            _cabi.check(_cabi.load().jp_halo_exchange(C.c_void_p(p._ctx), C.c_void_p(comm.handle if comm is not None else None), nbr,
                                                      _ptr_array(arrays), len(arrays), C.c_void_p(p.index.data_ptr()), _stream()),
                        "update_cell_halo")
            return sum(plane_bytes(p.ncells, p.max_xcell, d, len(arrays)) for d in range(p.ndim) for k in range(2)
                       if nbr[2 * d + k] not in (-1, topo.rank))
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
                        buffers: Optional[dict] = None, classify: Optional[bool] = None, join: bool = True,
                        comm: Optional[Comm] = None) -> int:
    """``advection!(particles, method, V, dt)`` followed by ``update_cell_halo!(coords..., args..., index)``
    (scripts/temperature_advection3D_MPI.jl:83-91) with the exchange hidden behind the advection: the bricks holding the two
    outermost cell layers are advected first (``jp_advect_region`` SHELL), the planes are then packed, sent over NCCL and unpacked
    on a high-priority side stream while the current stream advects the interior bricks; the current stream waits for the side
    stream before returning control to the next call (``move_particles``).  Same results as the two calls in sequence."""
    if topo is None or not topo.decomposed:
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
            sent = update_cell_halo(p, args, topo, group=group, buffers=buffers, comm=comm)
            done = torch.cuda.Event()
            done.record(side)
        advection(p, method, V, dt, region="interior")
        if join:
            main.wait_event(done)
        else:
            _PENDING[id(p)] = done          # the caller joins with join_halo(p) (bench.py: to time what is left of the exchange)
    return sent
