"""Host-side mirror of JustPIC.jl's public API for the particle-in-cell hot path.

Same names (minus Julia's ``!``), argument order and error behaviour as the
reference, so a test written against the reference reads the same here:

    particles = init_particles(CUDABackend, 24, 48, 12, grid_vx, grid_vy, grid_vz)
    pT, = init_cell_arrays(particles, 1)
    grid2particle(pT, T, particles)
    advection(particles, RungeKutta2(), V, dt)
    move_particles(particles, (pT,))
    inject_particles(particles, (pT,))
    particle2grid(T, pT, particles)

Every call goes through the C ABI of libjustpic_sm100a.so (include/justpic_c.h);
torch is only the allocator / stream provider.  There is no CPU path.

Memory layout.  A CellArray with S slots over cells (nx, ny[, nz]) is a
contiguous torch tensor of shape (S, [nz,] ny, nx): element (slot s, cell
i,j,k) sits at i + nx*(j + ny*k) + s*C, i.e. exactly the reference's CUDA
CellArray layout (ext/JustPICCUDAExt.jl:26-30).  Grid fields are torch tensors
of shape ([nz,] ny, nx) (+1 for vertex fields) = Julia's column-major
(nx, ny[, nz]) arrays.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi

__all__ = [
    "CUDABackend", "LinRange", "StepRange", "expand_range", "add_ghost_nodes",
    "Euler", "RungeKutta2", "RungeKutta4",
    "Particles", "PhaseRatios",
    "init_particles", "init_cell_arrays", "cell_array",
    "advection", "advection_LinP", "advection_MQS", "move_particles", "inject_particles", "inject_particles_phase", "force_injection", "clean_particles",
    "grid2particle", "grid2particle_flip", "centroid2particle", "particle2grid", "particle2centroid",
    "SubgridDiffusionCellArrays", "subgrid_diffusion", "subgrid_diffusion_centroid",
    "phase_ratios_center", "phase_ratios_vertex", "phase_ratios_face", "phase_ratios_midpoint",
    "update_phase_ratios", "set_synchronous",
    "Array", "CuArray", "HostParticles", "HostPhaseRatios", "last_move_classify",
    "move_interp_handoff", "last_interp_handoff", "invalidate_handoffs", "profile_move", "read_move_profile",
    "capture_step", "graph_step_offset",
]


class CUDABackend:
    """Backend tag (reference: CUDA.CUDABackend).  The only backend there is."""


_SYNC = False
MOVE_MODE = "auto"      # module default for move_particles
MOVE_POLICY = "reference"   # module default slot policy ("compact" is an opt-in deviation from the reference)
P2G_MODE = "twopass_fastw"   # module default for particle2grid (see its docstring)
PHASE_MODE = "fused"         # module default for update_phase_ratios ("literal" = the reference's kernels, bit-exact)


def set_synchronous(flag: bool) -> None:
    """``True`` reproduces the reference's launch-then-synchronize semantics
    (src/launch.jl:60-69); default is stream-ordered asynchronous execution."""
    global _SYNC
    _SYNC = bool(flag)


# --------------------------------------------------------------------------- grids
class LinRange:
    """Julia ``LinRange(start, stop, len)``: element i (0-based) is
    ``(1-t)*start + t*stop`` with ``t = i/(len-1)`` (Base.lerpi).  Passing
    LinRange grids selects the reference's *range* code path (scalar spacings
    ``x[2]-x[1]``, src/Particles/particles_utils.jl:108-166); passing arrays
    selects the vector path (``diff(x)``, :48-106)."""

    def __init__(self, start: float, stop: float, length: int):
        self.start, self.stop, self.len = float(start), float(stop), int(length)

    def __len__(self) -> int:
        return self.len

    def __array__(self, dtype=None, copy=None):
        j = np.arange(self.len, dtype=np.float64)
        t = j / float(self.len - 1) if self.len > 1 else j
        a = (1.0 - t) * self.start + t * self.stop
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, i):
        return np.asarray(self)[i]


class StepRange(LinRange):
    """Julia ``range(start, stop, length=len)`` (``StepRangeLen`` with twice-precision arithmetic):
    element i is the correctly rounded value of ``start + i*(stop-start)/(len-1)``, e.g. exact
    multiples of 2^-k on power-of-two grids (the README's grids, README.md:24-48), whereas
    ``LinRange`` rounds ``(1-t)*start + t*stop`` term by term.  Same *range* code path as LinRange."""

    def __array__(self, dtype=None, copy=None):
        from fractions import Fraction
        a, b, d = Fraction(self.start), Fraction(self.stop), max(self.len - 1, 1)
        out = np.array([float(a + (b - a) * i / d) for i in range(self.len)], dtype=np.float64)
        return out if dtype is None else out.astype(dtype)


def expand_range(x: LinRange) -> LinRange:
    """``expand_range`` of the reference scripts/tests (e.g.
    scripts/temperature_advection3D.jl:12-19): one ghost node on either side."""
    a = np.asarray(x)
    dx = a[1] - a[0]
    return type(x)(a.min() - dx, a.max() + dx, len(a) + 2)


def add_ghost_nodes(x, dx, origin=None) -> np.ndarray:
    """``add_ghost_nodes`` (src/Utils.jl:10-16) for array grids."""
    a = np.asarray(x, dtype=np.float64)
    return np.concatenate(([a.min() - dx], a, [a.max() + dx]))


# --------------------------------------------------------------------------- integrators
class Euler:
    """``Euler()`` (src/Advection/types.jl:17-19); accepts and ignores any arguments."""
    scheme = 0
    alpha = 0.0

    def __init__(self, *_):
        pass


class RungeKutta2:
    """``RungeKutta2(α=0.5)`` (src/Advection/types.jl:27-40): requires 0 < α < 1."""
    scheme = 1

    def __init__(self, alpha: float = 0.5):
        if not (0 < alpha < 1):
            raise ValueError("Only 0 < α < 1 is supported")
        self.alpha = float(alpha)


class RungeKutta4:
    """``RungeKutta4()`` (src/Advection/types.jl:47-49)."""
    scheme = 2
    alpha = 0.0

    def __init__(self, *_):
        pass


# --------------------------------------------------------------------------- containers
def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _done() -> None:
    if _SYNC:
        torch.cuda.current_stream().synchronize()


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (C.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


@dataclass
class Particles:
    """Mirror of the reference ``Particles`` struct (src/particles.jl:17-46)."""
    coords: Tuple[torch.Tensor, ...]
    index: torch.Tensor
    nxcell: int
    max_xcell: int
    min_xcell: int
    np: int
    di: SimpleNamespace
    xci: Tuple[np.ndarray, ...]
    xvi: Tuple[np.ndarray, ...]
    xi_vel: Tuple[Tuple[np.ndarray, ...], ...]
    uniform: bool
    seed: int = 42
    _ctx: Optional[int] = field(default=None, repr=False)
    _keep: list = field(default_factory=list, repr=False)
    _inject_step: int = 0

    @property
    def ndim(self) -> int:
        return len(self.coords)

    @property
    def ncells(self) -> Tuple[int, ...]:
        """(nx, ny[, nz])"""
        return tuple(len(x) for x in self.xci)

    @property
    def device(self) -> torch.device:
        return self.index.device

    def _c(self) -> _cabi.ParticlesC:
        pc = _cabi.ParticlesC()
        for d in range(3):
            pc.coords[d] = self.coords[d].data_ptr() if d < self.ndim else None
        pc.index = self.index.data_ptr()
        return pc

    def __del__(self):
        try:
            if self._ctx:
                _cabi.load().jp_ctx_destroy(C.c_void_p(self._ctx))
                self._ctx = None
        except Exception:
            pass


def _make_ctx(ndim, n, S, uniform, xvi, xci, xi_vel, device_index) -> Tuple[int, list]:
    lib = _cabi.load()
    keep = []
    gd = _cabi.GridDesc()
    gd.ndim, gd.S, gd.uniform = ndim, S, 1 if uniform else 0

    def ptr(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return a.ctypes.data_as(_cabi.c_double_p)

    for d in range(3):
        gd.n[d] = n[d] if d < ndim else 1
    for d in range(ndim):
        gd.xv[d] = ptr(xvi[d])
        gd.xc[d] = ptr(xci[d])
        for c in range(ndim):
            gd.xvel[c][d] = ptr(xi_vel[c][d])
            gd.nvel[c][d] = len(xi_vel[c][d])
    out = C.c_void_p()
    _cabi.check(lib.jp_ctx_create(C.byref(gd), device_index, C.byref(out)), "jp_ctx_create")
    return out.value, keep


def cell_array(x, ncells: Sequence[int], ni: Sequence[int], device=None, dtype=None) -> torch.Tensor:
    """``cell_array(backend, x, ncells, ni)`` (src/launch.jl:101-108): a CellArray with
    ``prod(ncells)`` entries per cell over the grid ``ni`` = (nx, ny[, nz]), filled with ``x``."""
    if dtype is None:
        dtype = torch.uint8 if isinstance(x, (bool, np.bool_)) else torch.float64
    shape = (int(np.prod(ncells)), *reversed([int(v) for v in ni]))
    return torch.full(shape, x, dtype=dtype, device=device or "cuda")


def init_particles(backend, nxcell: int, max_xcell: int, min_xcell: int, *xi_vel, seed: int = 42,
                   device=None) -> Particles:
    """``init_particles(backend, nxcell, max_xcell, min_xcell, grid_vx, grid_vy[, grid_vz])``
    (src/Particles/particles_utils.jl:30-35, :48-166): random-in-quadrant seeding.
    ``xi_vel[i]`` is the tuple of 1-D coordinate vectors of velocity component i
    (LinRange -> range path, arrays -> vector path)."""
    if len(xi_vel) == 1 and isinstance(xi_vel[0], (tuple, list)) and len(xi_vel[0]) and isinstance(xi_vel[0][0], (tuple, list)):
        xi_vel = tuple(xi_vel[0])
    if len(xi_vel) == 0:
        raise ValueError("The velocity grid cannot be empty")
    N = len(xi_vel)
    if N not in (2, 3) or any(len(g) != N for g in xi_vel):
        raise ValueError("expected N velocity grids of N coordinate vectors each, N = 2 or 3")
    if not isinstance(nxcell, (int, np.integer)):
        raise NotImplementedError("regular (NTuple nxcell) particle layouts are dead code in the reference and not supported")
    uniform = all(isinstance(x, LinRange) for g in xi_vel for x in g)
    xv_np = tuple(tuple(np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in g) for g in xi_vel)
    # centre / vertex grids exactly as the reference derives them (:56-74 / :117-135)
    if N == 3:
        xci = (xv_np[1][0][1:-1].copy(), xv_np[0][1][1:-1].copy(), xv_np[0][2][1:-1].copy())
    else:
        xci = (xv_np[1][0][1:-1].copy(), xv_np[0][1][1:-1].copy())
    xvi = tuple(xv_np[i][i] for i in range(N))
    if uniform:
        di = SimpleNamespace(center=tuple(x[1] - x[0] for x in xci), vertex=tuple(x[1] - x[0] for x in xvi),
                             velocity=tuple(tuple(x[1] - x[0] for x in g) for g in xv_np))
    else:
        di = SimpleNamespace(center=tuple(np.diff(x) for x in xci), vertex=tuple(np.diff(x) for x in xvi),
                             velocity=tuple(tuple(np.diff(x) for x in g) for g in xv_np))
    ni = tuple(len(x) for x in xci)
    NQ = 4 if N == 2 else 8
    np_quadrant = math.ceil(nxcell / NQ)
    nxcell = np_quadrant * NQ
    max_xcell = max(nxcell, int(max_xcell))
    npart = max_xcell * int(np.prod(ni))
    dev = torch.device(device or "cuda")
    if dev.type != "cuda":
        raise RuntimeError("justpic.jl_b200 has no CPU path: particles must live on a CUDA device")
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    dev = torch.device("cuda", dev_index)
    ctx, keep = _make_ctx(N, ni, max_xcell, uniform, xvi, xci, xv_np, dev_index)
    with torch.cuda.device(dev):
        coords = tuple(torch.empty((max_xcell, *reversed(ni)), dtype=torch.float64, device=dev) for _ in range(N))
        index = torch.empty((max_xcell, *reversed(ni)), dtype=torch.uint8, device=dev)
        p = Particles(coords, index, nxcell, max_xcell, int(min_xcell), npart, di, xci, xvi, xv_np, uniform,
                      seed=int(seed), _ctx=ctx, _keep=keep)
        pc = p._c()
        _cabi.check(_cabi.load().jp_init_particles(C.c_void_p(ctx), C.byref(pc), nxcell, C.c_uint64(int(seed)), _stream()),
                    "init_particles")
        _done()
    return p


def init_cell_arrays(particles: Particles, n: int) -> Tuple[torch.Tensor, ...]:
    """``init_cell_arrays(particles, Val(N))`` (src/Particles/particles_utils.jl:294-301)."""
    return tuple(torch.zeros_like(particles.coords[0]) for _ in range(int(n)))


def _field(t: torch.Tensor, p: Particles, numel: int, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or t.dtype != torch.float64 or t.device != p.device or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous float64 tensor on {p.device}")
    if t.numel() != numel:
        raise ValueError(f"{name}: expected {numel} elements, got {t.numel()}")
    return t


def _pfield(t, p, name):
    return _field(t, p, p.index.numel(), name)


def _nodes(p: Particles, plus: int) -> int:
    return int(np.prod([n + plus for n in p.ncells]))


def _args(args, p: Particles):
    if isinstance(args, torch.Tensor):
        args = (args,)
    args = tuple(args)
    if len(args) > _cabi.JP_MAX_ARGS:
        raise ValueError(f"at most {_cabi.JP_MAX_ARGS} particle fields per call")
    for i, a in enumerate(args):
        _pfield(a, p, f"args[{i}]")
    return args


# --------------------------------------------------------------------------- hot path
def advection(particles: Particles, method, V, dt: float, affine: Optional[bool] = None,
              classify: Optional[bool] = None, region: Optional[str] = None) -> None:
    """``advection!(particles, method, V, dt)`` (src/Particles/Advection/advection.jl:21-62).
    ``affine=False`` forces grid-coordinate table look-ups even when the grid vectors were verified
    to be exactly affine (identical results; the parity tests run both).
    ``classify=True`` switches on the advection -> move hand-off (JP_OPT_ADVECT_CLASSIFY in
    include/justpic_c.h; sticky per ``Particles``): the kernel also classifies every new position for the
    following ``move_particles``, which then skips its own pass over the coordinates.  Results are
    bit-identical; the caller must not write coordinates / index between the two calls by other means.
    ``region="shell"`` then ``region="interior"``: the same advection in two launches (jp_advect_region) -- first every
    brick of cells holding one of the two outermost cell layers, then the rest -- so that ``update_cell_halo`` can run
    on another stream while the interior is advected (``halo.advection_with_halo``)."""
    p = particles
    if classify is not None:
        _cabi.check(_cabi.load().jp_set_option(C.c_void_p(p._ctx), _cabi.JP_OPT_ADVECT_CLASSIFY, 1 if classify else 0), "jp_set_option")
    if affine is not None:
        _cabi.check(_cabi.load().jp_set_option(C.c_void_p(p._ctx), _cabi.JP_OPT_ADVECT_AFFINE, 1 if affine else 0), "jp_set_option")
    V = tuple(V)
    if len(V) != p.ndim:
        raise ValueError("V must hold one staggered array per dimension")
    for c, v in enumerate(V):
        _field(v, p, int(np.prod([len(x) for x in p.xi_vel[c]])), f"V[{c}]")
    lib = _cabi.load()
    pc = p._c()
    with torch.cuda.device(p.device):
        regions = {None: _cabi.JP_REGION_ALL, "all": _cabi.JP_REGION_ALL, "shell": _cabi.JP_REGION_SHELL, "interior": _cabi.JP_REGION_INTERIOR}
        if region not in regions:
            raise ValueError("advection: region must be None, 'shell' or 'interior'")
        _cabi.check(lib.jp_advect_region(C.c_void_p(p._ctx), C.byref(pc), method.scheme, float(method.alpha),
                                         _ptr_array(V), float(dt), regions[region], _stream()), "advection")
        _done()


def _advection_interp(particles: Particles, method, V, dt: float, interp: int, who: str) -> None:
    p = particles
    V = tuple(V)
    if len(V) != p.ndim:
        raise ValueError("V must hold one staggered array per dimension")
    for c, v in enumerate(V):
        _field(v, p, int(np.prod([len(x) for x in p.xi_vel[c]])), f"V[{c}]")
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_advect_interp(C.c_void_p(p._ctx), C.byref(pc), method.scheme, float(method.alpha),
                                                  _ptr_array(V), float(dt), interp, _stream()), who)
        _done()


def advection_LinP(particles: Particles, method, V, dt: float) -> None:
    """``advection_LinP!(particles, method, V, dt)`` (src/Particles/Advection/advection_LinP.jl:12-26)."""
    _advection_interp(particles, method, V, dt, 1, "advection_LinP")


def advection_MQS(particles: Particles, method, V, dt: float) -> None:
    """``advection_MQS!(particles, method, V, dt)`` (src/Particles/Advection/advection_MQS.jl:16-29)."""
    _advection_interp(particles, method, V, dt, 2, "advection_MQS")


def advect_affine_level(particles: Particles) -> int:
    """0: the tiled advection kernel looks grid coordinates up; 1: it regenerates vertex coordinates
    arithmetically; 2: ghosted-centre coordinates too (vectors verified exactly affine at context
    creation and the option not switched off)."""
    v = C.c_int32(0)
    _cabi.check(_cabi.load().jp_get_option(C.c_void_p(particles._ctx), _cabi.JP_OPT_ADVECT_AFFINE, C.byref(v)), "jp_get_option")
    return int(v.value)


_MOVE_POLICIES = {"reference": _cabi.JP_MOVE_POLICY_REFERENCE, "compact": _cabi.JP_MOVE_POLICY_COMPACT, "dense": _cabi.JP_MOVE_POLICY_DENSE}


def move_particles(particles: Particles, args=(), mode: Optional[str] = None, policy: Optional[str] = None) -> None:
    """``move_particles!(particles, args)`` (src/Particles/move_safe.jl:21-49).
    ``policy``: "reference" (default: the reference's free-slot rule, bit for bit) or "compact" (opt-in, NOT
    reference behaviour: every migrant takes the lowest free slot of its destination, which keeps slot planes
    dense) or "dense" (opt-in, NOT reference behaviour: all leavers vacate first, then every migrant takes the lowest
    free slot; planned path only); see JP_OPT_MOVE_POLICY in include/justpic_c.h.
    ``mode``: "auto" (default: sweeps planned on occupancy words + streaming payload passes,
    direct sweeps when a particle sits exactly on a face) or "direct" (always the literal
    sweeps on the particle arrays).  Both give the reference's slot assignment bit for bit."""
    p = particles
    args = _args(args, p)
    pc = p._c()
    m = (mode or MOVE_MODE).lower()
    if m not in ("auto", "direct"):
        raise ValueError("move_particles mode must be 'auto' or 'direct'")
    pol = (policy or MOVE_POLICY).lower()
    if pol not in _MOVE_POLICIES:
        raise ValueError("move_particles policy must be 'reference', 'compact' or 'dense'")
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_set_option(C.c_void_p(p._ctx), _cabi.JP_OPT_MOVE_POLICY,
                                               _MOVE_POLICIES[pol]), "jp_set_option")
        _cabi.check(_cabi.load().jp_set_option(C.c_void_p(p._ctx), _cabi.JP_OPT_MOVE_MODE,
                                               _cabi.JP_MOVE_AUTO if m == "auto" else _cabi.JP_MOVE_DIRECT), "jp_set_option")
        _cabi.check(_cabi.load().jp_move(C.c_void_p(p._ctx), C.byref(pc), _ptr_array(args), len(args), _stream()),
                    "move_particles")
        _done()


def move_stats(particles: Particles) -> Tuple[int, int, int]:
    """(moved, dropped, deleted) of the last ``move_particles`` (synchronises)."""
    out = (C.c_int64 * 3)()
    with torch.cuda.device(particles.device):
        _cabi.check(_cabi.load().jp_move_stats(C.c_void_p(particles._ctx), out, _stream()), "move_stats")
    return int(out[0]), int(out[1]), int(out[2])


def last_move_path(particles: Particles) -> str:
    """"plan" or "direct": which implementation the last ``move_particles`` call took."""
    return "plan" if (_cabi.load().jp_last_move_path(C.c_void_p(particles._ctx)) & 0xff) == 0 else "direct"


def last_move_classify(particles: Particles) -> str:
    """"handoff" if the last planned ``move_particles`` built its plan from the words left by
    ``advection(..., classify=True)``, "coords" if it classified the coordinates itself."""
    v = C.c_int32(0)
    _cabi.check(_cabi.load().jp_get_option(C.c_void_p(particles._ctx), _cabi.JP_OPT_LAST_CLASSIFY, C.byref(v)), "jp_get_option")
    return "handoff" if v.value else "coords"


def move_interp_handoff(particles: Particles, Fp: Optional[torch.Tensor] = None, phases: Optional[torch.Tensor] = None,
                        nphases: int = 0, enable: bool = True) -> None:
    """Move -> interpolation hand-off (JP_OPT_MOVE_INTERP in include/justpic_c.h; sticky per ``Particles``, opt-in).
    Tell the library which particle field the ``particle2grid(F, Fp, particles)`` after the next ``move_particles`` will
    interpolate and which field / how many phases ``phase_ratios_center(phase_ratios, particles, phases)`` will use: the last
    streaming pass of ``move_particles`` then also accumulates that call's per-cell sums / the centre ratios (same arithmetic,
    same order: bit-identical results), and the two calls do not read the particles again.  Both fields must be among the
    ``args`` of ``move_particles``; the caller must not write the particle arrays or the two fields between ``move_particles``
    and the consumers other than through this API (which drops the hand-off where needed)."""
    p = particles
    lib = _cabi.load()
    fp = _pfield(Fp, p, "Fp").data_ptr() if Fp is not None else None
    ph = _pfield(phases, p, "phases").data_ptr() if phases is not None else None
    _cabi.check(lib.jp_set_option(C.c_void_p(p._ctx), _cabi.JP_OPT_MOVE_INTERP, 1 if enable else 0), "jp_set_option")
    _cabi.check(lib.jp_move_interp_fields(C.c_void_p(p._ctx), C.c_void_p(fp), C.c_void_p(ph), int(nphases) if ph else 0),
                "move_interp_handoff")


def invalidate_handoffs(particles: Particles) -> None:
    """Call after writing ``particles.coords`` / ``particles.index`` / a registered particle field by other means than this API
    between ``advection`` and ``move_particles`` (or ``move_particles`` and its consumers): drops what the hand-offs left."""
    _cabi.check(_cabi.load().jp_invalidate_handoffs(C.c_void_p(particles._ctx)), "jp_invalidate_handoffs")


def last_interp_handoff(particles: Particles) -> Tuple[bool, bool]:
    """(particle2grid, phase_ratios_center): whether the last call of each used the sums left by ``move_particles``."""
    v = C.c_int32(0)
    _cabi.check(_cabi.load().jp_get_option(C.c_void_p(particles._ctx), _cabi.JP_OPT_LAST_INTERP, C.byref(v)), "jp_get_option")
    return bool(v.value & 1), bool(v.value & 2)


def graph_step_offset(particles: Particles, value: Optional[int] = None) -> int:
    """JP_OPT_GRAPH_STEP_OFFSET: the device word a captured ``inject_particles`` advances on every replay (replay i injects with
    ``step + i``).  ``value`` sets it (0 when going back to eager calls or capturing anew); returns the current value."""
    with torch.cuda.device(particles.device):
        if value is not None:
            _cabi.check(_cabi.load().jp_set_option(C.c_void_p(particles._ctx), _cabi.JP_OPT_GRAPH_STEP_OFFSET, int(value)), "jp_set_option")
        v = C.c_int32(0)
        _cabi.check(_cabi.load().jp_get_option(C.c_void_p(particles._ctx), _cabi.JP_OPT_GRAPH_STEP_OFFSET, C.byref(v)), "jp_get_option")
    return int(v.value)


def capture_step(particles: Particles, step_fn, warmup: int = 1) -> "torch.cuda.CUDAGraph":
    """Capture one call of ``step_fn()`` -- any sequence of this module's calls on ``particles`` with fixed tensors and scalars, e.g.
    advection + move_particles + inject_particles + particle2grid -- into a CUDA graph; ``graph.replay()`` then runs the step with
    one launch (what a launch-bound small grid wants; include/justpic_c.h, JP_OPT_GRAPH_STEP_OFFSET).  ``step_fn`` runs ``warmup``
    times eagerly first (that sizes every workspace; the library refuses to allocate while capturing).  Replay i of the graph is
    what the (warmup + 1 + i)-th eager call would have been: the arguments are frozen, tensors are read where they live (write new
    velocities into the same tensors), and ``inject_particles``' step counter advances on the device."""
    for _ in range(max(1, int(warmup))):
        step_fn()
    torch.cuda.synchronize(particles.device)
    graph_step_offset(particles, 0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.device(particles.device), torch.cuda.graph(g):
        step_fn()
    return g


def profile_move(particles: Particles, enable: bool = True) -> None:
    """JP_OPT_PROFILE: record CUDA events between the stages of the planned ``move_particles`` (no synchronisation)."""
    _cabi.check(_cabi.load().jp_set_option(C.c_void_p(particles._ctx), _cabi.JP_OPT_PROFILE, 1 if enable else 0), "jp_set_option")


def read_move_profile(particles: Particles) -> dict:
    """Mean ms per stage of the planned ``move_particles`` over the calls since the last read (synchronises the device)."""
    out = (C.c_double * 5)()
    n = C.c_int32(0)
    with torch.cuda.device(particles.device):
        _cabi.check(_cabi.load().jp_profile_read(C.c_void_p(particles._ctx), out, C.byref(n)), "jp_profile_read")
    return {"calls": int(n.value), "classify": out[0], "plan": out[1], "finalize_scan": out[2], "gather": out[3], "scatter": out[4]}


def last_move_reasons(particles: Particles) -> int:
    """Why the last move fell back to the direct sweeps (bit mask: 1 displacement > 1 cell, 2 particle on a
    face of its own cell, 4 particle on a face of its destination)."""
    return _cabi.load().jp_last_move_path(C.c_void_p(particles._ctx)) >> 8


def inject_particles(particles: Particles, args=(), step: Optional[int] = None) -> None:
    """``inject_particles!(particles, args)`` (src/Particles/injection.jl:19-53).  The reference
    draws from the backend's global RNG; here the stream is Philox keyed
    (particles.seed, step, cell, slot) with ``step`` auto-incremented per call."""
    p = particles
    args = _args(args, p)
    if step is None:
        step = p._inject_step
        p._inject_step += 1
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_inject(C.c_void_p(p._ctx), C.byref(pc), _ptr_array(args), len(args), p.min_xcell,
                                           C.c_uint64(p.seed), C.c_uint32(int(step)), _stream()), "inject_particles")
        _done()


def inject_particles_phase(particles: Particles, particles_phases: torch.Tensor, args=(), fields=(), step: Optional[int] = None) -> None:
    """``inject_particles_phase!(particles, particles_phases, args, fields)`` (src/Particles/injection.jl:146-200):
    new particles take the phase of their nearest neighbour and ``args[j]`` is interpolated from the grid
    field ``fields[j]`` (a centre field if it has one value per cell, else a vertex field), clamped to the
    extrema of the interpolation stencil.  RNG stream as in ``inject_particles``."""
    p = particles
    args = _args(args, p)
    fields = tuple(fields)
    if len(fields) != len(args):
        raise ValueError("inject_particles_phase: one grid field per particle field")
    ph = _pfield(particles_phases, p, "particles_phases")
    ncell, nvert = int(np.prod(p.ncells)), _nodes(p, 1)
    kinds = (C.c_int32 * max(len(args), 1))()
    for j, f in enumerate(fields):
        if not isinstance(f, torch.Tensor) or f.numel() not in (ncell, nvert):
            raise ValueError(f"fields[{j}]: expected a centre ({ncell}) or vertex ({nvert}) field")
        kinds[j] = 1 if f.numel() == ncell else 0
        _field(f, p, f.numel(), f"fields[{j}]")
    if step is None:
        step = p._inject_step
        p._inject_step += 1
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_inject_phase(C.c_void_p(p._ctx), C.byref(pc), C.c_void_p(ph.data_ptr()), _ptr_array(args),
                                                 _ptr_array(fields), kinds, len(args), p.min_xcell, C.c_uint64(p.seed),
                                                 C.c_uint32(int(step)), _stream()), "inject_particles_phase")
        _done()


def inject_stats(particles: Particles) -> int:
    out = C.c_int64()
    with torch.cuda.device(particles.device):
        _cabi.check(_cabi.load().jp_inject_stats(C.c_void_p(particles._ctx), C.byref(out), _stream()), "inject_stats")
    return int(out.value)


def force_injection(particles: Particles, p_new, fields=(), values=()) -> None:
    """``force_injection!(particles, p_new, fields, values)`` (src/Particles/forced_injection.jl:16-29): candidate points are
    written straight into free slots, no nearest-neighbour search.  ``p_new``: one tensor per coordinate, CellArray-shaped like
    ``particles.coords`` (entry k of cell I at ``[k, ..., I]``); a cell whose FIRST entry is NaN injects nothing ("NaN marks
    empty input slots"), otherwise every free slot ip takes entry ip.  ``fields[j]`` of the filled slots is set to ``values[j]``."""
    p = particles
    fields = _args(fields, p)
    values = tuple(float(v) for v in values)
    if len(values) != len(fields):
        raise ValueError("force_injection: one value per field")
    p_new = tuple(p_new)
    if len(p_new) != p.ndim:
        raise ValueError(f"force_injection: p_new needs {p.ndim} coordinate arrays")
    p_new = tuple(_pfield(t, p, f"p_new[{d}]") for d, t in enumerate(p_new))
    vals = (C.c_double * max(len(values), 1))(*values)
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_force_injection(C.c_void_p(p._ctx), C.byref(pc), _ptr_array(p_new), _ptr_array(fields), vals,
                                                    len(fields), _stream()), "force_injection")
        _done()


def clean_particles(particles: Particles, grid=None, args=()) -> None:
    """``clean_particles!(particles, grid, args)`` (src/Particles/move_safe.jl:289-320);
    ``grid`` must be the particles' own vertex grid (or None)."""
    p = particles
    args = _args(args, p)
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_clean(C.c_void_p(p._ctx), C.byref(pc), _ptr_array(args), len(args), _stream()),
                    "clean_particles")
        _done()


def _call5(fn_name, p, a, b, who):
    pc = p._c()
    with torch.cuda.device(p.device):
        fn = getattr(_cabi.load(), fn_name)
        _cabi.check(fn(C.c_void_p(p._ctx), C.byref(pc), C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), _stream()), who)
        _done()


def grid2particle(Fp, F, particles: Particles) -> None:
    """``grid2particle!(Fp, F, particles)`` (src/Interpolations/grid_to_particle.jl:26-35)."""
    _call5("jp_grid2particle", particles, _pfield(Fp, particles, "Fp"), _field(F, particles, _nodes(particles, 1), "F"),
           "grid2particle")


def grid2particle_flip(Fp, xvi, F, F0, particles: Particles, alpha: float = 0.0) -> None:
    """``grid2particle_flip!(Fp, xvi, F, F0, particles; α = 0.0)`` (src/Interpolations/grid_to_particle.jl:125-138):
    PIC/FLIP blend (α = 1 pure PIC, α = 0 pure FLIP).  ``xvi`` must be the particles' vertex grid (or None)."""
    p = particles
    Fp = _pfield(Fp, p, "Fp")
    F = _field(F, p, _nodes(p, 1), "F")
    F0 = _field(F0, p, _nodes(p, 1), "F0")
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_grid2particle_flip(C.c_void_p(p._ctx), C.byref(pc), C.c_void_p(Fp.data_ptr()),
                                                       C.c_void_p(F.data_ptr()), C.c_void_p(F0.data_ptr()), float(alpha), _stream()),
                    "grid2particle_flip")
        _done()


class SubgridDiffusionCellArrays:
    """``SubgridDiffusionCellArrays(particles; loc = :vertex)`` (src/Physics/subgrid_diffusion.jl:13-33): scratch
    CellArrays ``pT0``, ``pΔT`` (``pdT``), ``dt₀`` (``dt0``) and the grid-sized ``ΔT_subgrid`` (``dT_subgrid``)."""

    def __init__(self, particles: Particles, loc: str = "vertex"):
        loc = str(loc).lstrip(":")
        if loc not in ("vertex", "center"):
            raise ValueError("loc must be :vertex or :center")
        self.loc = loc
        self.pdT, self.pT0, self.dt0 = init_cell_arrays(particles, 3)
        plus = 1 if loc == "vertex" else 0
        self.dT_subgrid = torch.zeros(tuple(reversed([n + plus for n in particles.ncells])), dtype=torch.float64,
                                      device=particles.device)


def _subgrid(pT, T_grid, dT_grid, sa: SubgridDiffusionCellArrays, p: Particles, dt, d, centroid, who):
    plus = 0 if centroid else 1
    if sa.loc != ("center" if centroid else "vertex"):
        raise ValueError(f"{who}: subgrid arrays were allocated for loc = :{sa.loc}")
    pT = _pfield(pT, p, "pT")
    T_grid = _field(T_grid, p, _nodes(p, plus), "T_grid")
    if not isinstance(dT_grid, torch.Tensor) or dT_grid.dtype != torch.float64 or not dT_grid.is_contiguous() or dT_grid.dim() != p.ndim:
        raise ValueError("dT_grid: expected a contiguous float64 grid array")
    ext = (C.c_int32 * 3)(*(list(reversed(dT_grid.shape)) + [1] * (3 - p.ndim)))
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_subgrid_diffusion(
            C.c_void_p(p._ctx), C.byref(pc), C.c_void_p(pT.data_ptr()), C.c_void_p(T_grid.data_ptr()), C.c_void_p(dT_grid.data_ptr()),
            ext, C.c_void_p(sa.pT0.data_ptr()), C.c_void_p(sa.pdT.data_ptr()), C.c_void_p(sa.dt0.data_ptr()),
            C.c_void_p(sa.dT_subgrid.data_ptr()), float(dt), float(d), 1 if centroid else 0, _stream()), who)
        _done()


def subgrid_diffusion(pT, T_grid, dT_grid, subgrid_arrays, particles: Particles, dt: float, d: float = 1.0) -> None:
    """``subgrid_diffusion!(pT, T_grid, ΔT_grid, subgrid_arrays, particles, dt; d = 1.0)``
    (src/Physics/subgrid_diffusion.jl:55-78): vertex-grid subgrid diffusion correction of particle temperatures."""
    _subgrid(pT, T_grid, dT_grid, subgrid_arrays, particles, dt, d, False, "subgrid_diffusion")


def subgrid_diffusion_centroid(pT, T_grid, dT_grid, subgrid_arrays, particles: Particles, dt: float, d: float = 1.0) -> None:
    """``subgrid_diffusion_centroid!`` (src/Physics/subgrid_diffusion.jl:88-113): cell-centre variant."""
    _subgrid(pT, T_grid, dT_grid, subgrid_arrays, particles, dt, d, True, "subgrid_diffusion_centroid")


def centroid2particle(Fp, F, particles: Particles) -> None:
    """``centroid2particle!(Fp, F, particles)`` (src/Interpolations/centroid_to_particle.jl:13-20)."""
    _call5("jp_centroid2particle", particles, _pfield(Fp, particles, "Fp"), _field(F, particles, _nodes(particles, 0), "F"),
           "centroid2particle")


def particle2grid(F, Fp, particles: Particles, mode: Optional[str] = None) -> None:
    """``particle2grid!(F, Fp, particles)`` (src/Interpolations/particle_to_grid.jl:23-28).
    ``mode``: "twopass_fastw" (default: deterministic cell-partials + node gather with the weight
    evaluated as 1/sum(d^2); within the stated 1e-12 of the reference), "twopass" (same with the
    reference's inv(sqrt(.)^2) weight), or "exact" (the reference's single running sum: bit-exact,
    2^N x the traffic)."""
    m = (mode or P2G_MODE).lower()
    modes = {"exact": _cabi.JP_P2G_EXACT, "twopass": _cabi.JP_P2G_TWOPASS, "twopass_fastw": _cabi.JP_P2G_TWOPASS_FASTW}
    if m not in modes:
        raise ValueError("particle2grid mode must be one of " + ", ".join(modes))
    _cabi.check(_cabi.load().jp_set_option(C.c_void_p(particles._ctx), _cabi.JP_OPT_P2G_MODE, modes[m]), "jp_set_option")
    _call5("jp_particle2grid", particles, _field(F, particles, _nodes(particles, 1), "F"), _pfield(Fp, particles, "Fp"),
           "particle2grid")


def particle2centroid(F, Fp, particles: Particles) -> None:
    """``particle2centroid!(F, Fp, particles)`` (src/Interpolations/particle_to_grid_centroid.jl:10-16)."""
    _call5("jp_particle2centroid", particles, _field(F, particles, _nodes(particles, 0), "F"), _pfield(Fp, particles, "Fp"),
           "particle2centroid")


@dataclass
class PhaseRatios:
    """``PhaseRatios(backend, nphases, ni)`` (src/PhaseRatios/constructors.jl:18-47): CellArrays
    ``center`` (n), ``vertex`` (n+1), ``Vx/Vy[/Vz]`` (faces) and, in 3-D, ``yz/xz/xy`` (edge
    midpoints); in 2-D the last four are 1-cell dummies as in the reference."""

    def __init__(self, backend, nphases: int, ni: Sequence[int], device=None):
        ni = tuple(int(n) for n in ni)
        self.nphases = int(nphases)
        self.ni = ni
        mk = lambda dims: cell_array(0.0, (nphases,), dims, device=device)
        self.center = mk(ni)
        self.vertex = mk(tuple(n + 1 for n in ni))
        face = lambda d: tuple(n + (1 if i == d else 0) for i, n in enumerate(ni))
        self.Vx, self.Vy = mk(face(0)), mk(face(1))
        if len(ni) == 3:
            nx, ny, nz = ni
            self.Vz = mk(face(2))
            self.yz, self.xz, self.xy = mk((nx, ny + 1, nz + 1)), mk((nx + 1, ny, nz + 1)), mk((nx + 1, ny + 1, nz))
        else:
            self.Vz = self.yz = self.xz = self.xy = mk((1, 1))


_FACE_DIM = {"x": 0, "y": 1, "z": 2}
_MID_PLANE = {"xy": 0, "yz": 1, "xz": 2}


def _phase_call(fn_name, who, ratios, p, phases, K, nelem, *extra):
    r = _field(ratios, p, K * nelem, who)
    ph = _pfield(phases, p, "phases")
    pc = p._c()
    with torch.cuda.device(p.device):
        fn = getattr(_cabi.load(), fn_name)
        _cabi.check(fn(C.c_void_p(p._ctx), C.byref(pc), C.c_void_p(r.data_ptr()), C.c_void_p(ph.data_ptr()), K,
                       *extra, _stream()), who)
        _done()


def phase_ratios_vertex(phase_ratios: PhaseRatios, particles: Particles, phases: torch.Tensor) -> None:
    """``phase_ratios_vertex!(phase_ratios, particles, phases)`` (src/PhaseRatios/vertices.jl:4-13)."""
    p = particles
    _phase_call("jp_phase_ratios_vertex", "phase_ratios_vertex", phase_ratios.vertex, p, phases, phase_ratios.nphases,
                int(np.prod([n + 1 for n in p.ncells])))


def phase_ratios_face(phase_face: torch.Tensor, particles: Particles, phases: torch.Tensor, dimension: str) -> None:
    """``phase_ratios_face!(phase_face, particles, phases, dimension)`` (src/PhaseRatios/midpoints.jl:3-24);
    ``dimension`` is ``"x"``, ``"y"`` or ``"z"`` (the reference's ``:x/:y/:z``)."""
    p = particles
    d = _FACE_DIM.get(str(dimension).lstrip(":"))
    if d is None or d >= p.ndim:
        raise ValueError("dimension must be :x, :y or :z")
    K = int(phase_face.shape[0])
    nelem = int(np.prod([n + (1 if i == d else 0) for i, n in enumerate(p.ncells)]))
    _phase_call("jp_phase_ratios_face", "phase_ratios_face", phase_face, p, phases, K, nelem, d)


def phase_ratios_midpoint(phase_midpoint: torch.Tensor, particles: Particles, phases: torch.Tensor, dimension: str) -> None:
    """``phase_ratios_midpoint!(phase_midpoint, particles, phases, dimension)`` (src/PhaseRatios/midpoints.jl:115-126),
    3-D only; ``dimension`` is ``"xy"``, ``"yz"`` or ``"xz"``."""
    p = particles
    pl = _MID_PLANE.get(str(dimension).lstrip(":"))
    if pl is None or p.ndim != 3:
        raise ValueError("Unknown dimensions. Valid dimensions are :xy, :yz, :xz")
    off = {0: (1, 1, 0), 1: (0, 1, 1), 2: (1, 0, 1)}[pl]
    K = int(phase_midpoint.shape[0])
    nelem = int(np.prod([n + o for n, o in zip(p.ncells, off)]))
    _phase_call("jp_phase_ratios_midpoint", "phase_ratios_midpoint", phase_midpoint, p, phases, K, nelem, pl)


def update_phase_ratios(phase_ratios: PhaseRatios, particles: Particles, phases: torch.Tensor, mode: Optional[str] = None) -> None:
    """``update_phase_ratios!(phase_ratios, particles, phases)`` (src/PhaseRatios/utils.jl:15-41): centres,
    vertices, velocity nodes and (3-D) edge midpoints.  ``mode``: "literal" (the reference's kernels one after the
    other, bit-exact) or "fused" (default for up to 4 phases: one pass over the particles + node gathers, within the
    stated 1e-12; ~10x faster)."""
    pr, p = phase_ratios, particles
    m = (mode or PHASE_MODE).lower()
    if m not in ("literal", "fused"):
        raise ValueError("update_phase_ratios mode must be 'literal' or 'fused'")
    K = pr.nphases
    ph = _pfield(phases, p, "phases")
    _field(pr.center, p, K * int(np.prod(p.ncells)), "phase_ratios.center")
    _field(pr.vertex, p, K * _nodes(p, 1), "phase_ratios.vertex")
    faces = (pr.Vx, pr.Vy, pr.Vz)[:p.ndim]
    for d, f in enumerate(faces):
        _field(f, p, K * int(np.prod([n + (1 if i == d else 0) for i, n in enumerate(p.ncells)])), "phase_ratios face field")
    mids = (pr.xy, pr.yz, pr.xz) if p.ndim == 3 else ()
    for off, f in zip(((1, 1, 0), (0, 1, 1), (1, 0, 1)), mids):
        _field(f, p, K * int(np.prod([n + o for n, o in zip(p.ncells, off)])), "phase_ratios midpoint field")
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_update_phase_ratios(
            C.c_void_p(p._ctx), C.byref(pc), C.c_void_p(ph.data_ptr()), K, C.c_void_p(pr.center.data_ptr()),
            C.c_void_p(pr.vertex.data_ptr()), _ptr_array(faces), _ptr_array(mids) if mids else None,
            1 if m == "fused" else 0, _stream()), "update_phase_ratios")
        _done()


def phase_ratios_center(phase_ratios: PhaseRatios, particles: Particles, phases: torch.Tensor) -> None:
    """``phase_ratios_center!(phase_ratios, particles, phases)`` (src/PhaseRatios/centers.jl:3-30)."""
    p = particles
    K = phase_ratios.nphases
    r = _field(phase_ratios.center, p, K * int(np.prod(p.ncells)), "phase_ratios.center")
    ph = _pfield(phases, p, "phases")
    pc = p._c()
    with torch.cuda.device(p.device):
        _cabi.check(_cabi.load().jp_phase_ratios_center(C.c_void_p(p._ctx), C.byref(pc), C.c_void_p(r.data_ptr()),
                                                        C.c_void_p(ph.data_ptr()), K, _stream()), "phase_ratios_center")
        _done()


# --------------------------------------------------------------------------- Array(...) / CuArray(...)
# Host images of the device containers, in the reference's CPU CellArray layout
# (blocklength 1, data[1, S, C]: slot fastest; src/launch.jl:81): numpy arrays of shape
# ([nz,] ny, nx, S) in C order.  Checkpoints written from them are what JLD2 stores in the reference
# (src/IO/JLD2.jl saves Array(particles), test/test_save_load.jl:120-173).
@dataclass
class HostParticles:
    """``Array(particles)``: the ``Particles{CPU}`` image of a device container (src/CellArrays/conversion.jl:45-69)."""
    coords: Tuple[np.ndarray, ...]
    index: np.ndarray
    nxcell: int
    max_xcell: int
    min_xcell: int
    np: int
    di: SimpleNamespace
    xci: Tuple[np.ndarray, ...]
    xvi: Tuple[np.ndarray, ...]
    xi_vel: Tuple[Tuple[np.ndarray, ...], ...]
    uniform: bool
    seed: int = 42
    inject_step: int = 0


@dataclass
class HostPhaseRatios:
    """``Array(phase_ratios)``: every field of ``PhaseRatios`` in the host layout."""
    nphases: int
    ni: Tuple[int, ...]
    center: np.ndarray
    vertex: np.ndarray
    Vx: np.ndarray
    Vy: np.ndarray
    Vz: np.ndarray
    yz: np.ndarray
    xz: np.ndarray
    xy: np.ndarray


_PR_FIELDS = ("center", "vertex", "Vx", "Vy", "Vz", "yz", "xz", "xy")
_TORCH_OF = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}
_JP_OF = {torch.float64: _cabi.JP_F64, torch.float32: _cabi.JP_F32, torch.uint8: _cabi.JP_BOOL, torch.bool: _cabi.JP_BOOL}


def _eltype(T, src_dtype: torch.dtype) -> torch.dtype:
    if src_dtype in (torch.uint8, torch.bool):
        return src_dtype                                   # index stays Bool whatever T is (conversion.jl:50-51)
    if T is None:
        return src_dtype
    dt = np.dtype(T)
    if dt not in _TORCH_OF:
        raise TypeError(f"{T} is not a supported CellArray element type (Float64 / Float32)")
    return _TORCH_OF[dt]


def _permute(src: torch.Tensor, dst: torch.Tensor, ncells: int, ncomp: int, direction: int, ctx=None) -> None:
    with torch.cuda.device(src.device):
        _cabi.check(_cabi.load().jp_cellarray_permute(C.c_void_p(ctx), C.c_void_p(src.data_ptr()), _JP_OF[src.dtype],
                                                      C.c_void_p(dst.data_ptr()), _JP_OF[dst.dtype], ncells, ncomp, direction,
                                                      _stream()), "jp_cellarray_permute")


def _cellarray_to_host(t: torch.Tensor, T=None, ctx=None) -> np.ndarray:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and t.dim() >= 2):
        raise TypeError("Array(CellArray): expected a contiguous CUDA tensor of shape (ncomp, [nz,] ny, nx)")
    if t.dtype not in _JP_OF:
        raise TypeError(f"{t.dtype} is not a supported CellArray type.")
    ncomp, cells = int(t.shape[0]), tuple(int(v) for v in t.shape[1:])
    out = torch.empty((*cells, ncomp), dtype=_eltype(T, t.dtype), device=t.device)
    _permute(t, out, int(np.prod(cells)), ncomp, _cabi.JP_LAYOUT_TO_HOST, ctx)
    h = out.cpu().numpy()                                   # synchronises
    return h.astype(np.bool_) if t.dtype in (torch.uint8, torch.bool) else h


def _cellarray_to_device(a: np.ndarray, T=None, device=None, ctx=None) -> torch.Tensor:
    a = np.ascontiguousarray(a)
    if a.ndim < 2:
        raise TypeError("CuArray(CellArray): expected a host array of shape ([nz,] ny, nx, ncomp)")
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    elif a.dtype not in _TORCH_OF:
        raise TypeError(f"{a.dtype} is not a supported CellArray type.")
    dev = torch.device(device or "cuda")
    if dev.type != "cuda":
        raise RuntimeError("justpic.jl_b200 has no CPU path: CuArray targets a CUDA device")
    src = torch.from_numpy(a).to(dev)
    cells, ncomp = tuple(a.shape[:-1]), int(a.shape[-1])
    out = torch.empty((ncomp, *cells), dtype=_eltype(T, src.dtype), device=src.device)
    _permute(src, out, int(np.prod(cells)), ncomp, _cabi.JP_LAYOUT_TO_DEVICE, ctx)
    _done()
    return out


def Array(x, T=None):
    """``Array(x)`` / ``Array(T, x)`` (src/CellArrays/conversion.jl:19-69) for a device CellArray (tensor of
    shape ``(ncomp, [nz,] ny, nx)``), ``Particles`` or ``PhaseRatios``: the host image in the reference's CPU
    CellArray layout (``permutedims(data, (3, 2, 1))``: shape ``([nz,] ny, nx, ncomp)``), optionally
    converted to element type ``T`` (``index`` stays Bool).  The permutation runs on the device
    (``jp_cellarray_permute``), the copy to the host is one contiguous transfer per array."""
    if isinstance(x, (HostParticles, HostPhaseRatios, np.ndarray)):
        return x                                            # already on the host (conversion.jl:24-29)
    if isinstance(x, torch.Tensor):
        return _cellarray_to_host(x, T)
    if isinstance(x, Particles):
        return HostParticles(tuple(_cellarray_to_host(c, T, x._ctx) for c in x.coords), _cellarray_to_host(x.index, None, x._ctx),
                             x.nxcell, x.max_xcell, x.min_xcell, x.np, x.di, x.xci, x.xvi, x.xi_vel, x.uniform, x.seed,
                             x._inject_step)
    if isinstance(x, PhaseRatios):
        return HostPhaseRatios(x.nphases, x.ni, *(_cellarray_to_host(getattr(x, f), T) for f in _PR_FIELDS))
    raise TypeError(f"{type(x).__name__} is not a supported CellArray type.")


def CuArray(x, T=None, device=None):
    """``CuArray(x)`` / ``CuArray(T, x)`` (ext/JustPICCUDAExt.jl:51-187): the device container of a host image
    produced by :func:`Array` (or loaded from a checkpoint).  ``Particles`` are rebuilt with a fresh library
    context for the same grids; the kernels are fp64, so ``T`` other than Float64 is accepted for bare
    CellArrays and ``PhaseRatios`` only."""
    if isinstance(x, (Particles, PhaseRatios, torch.Tensor)):
        return x                                            # already on the device (ext/JustPICCUDAExt.jl:181-182)
    if isinstance(x, np.ndarray):
        return _cellarray_to_device(x, T, device)
    if isinstance(x, HostPhaseRatios):
        pr = PhaseRatios.__new__(PhaseRatios)
        pr.nphases, pr.ni = x.nphases, tuple(x.ni)
        for f in _PR_FIELDS:
            setattr(pr, f, _cellarray_to_device(getattr(x, f), T, device))
        return pr
    if isinstance(x, HostParticles):
        if T is not None and np.dtype(T) != np.float64:
            raise NotImplementedError("Particles on the device are Float64: the sm_100a kernels compute in fp64")
        dev = torch.device(device or "cuda")
        dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        dev = torch.device("cuda", dev_index)
        ni = tuple(len(c) for c in x.xci)
        ctx, keep = _make_ctx(len(ni), ni, x.max_xcell, x.uniform, x.xvi, x.xci, x.xi_vel, dev_index)
        coords = tuple(_cellarray_to_device(c.astype(np.float64, copy=False), None, dev, ctx) for c in x.coords)
        index = _cellarray_to_device(x.index, None, dev, ctx)
        return Particles(coords, index, x.nxcell, x.max_xcell, x.min_xcell, x.np, x.di, x.xci, x.xvi, x.xi_vel, x.uniform,
                         seed=x.seed, _ctx=ctx, _keep=keep, _inject_step=x.inject_step)
    raise TypeError(f"{type(x).__name__} is not a supported CellArray type.")
