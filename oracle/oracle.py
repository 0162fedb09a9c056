"""ctypes front-end of the CPU oracle (oracle/justpic_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product.
Arrays are numpy, CellArrays have shape (S, [nz,] ny, nx) (same memory layout
as the product's torch tensors).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libjustpic_oracle.so"
c_double_p = C.POINTER(C.c_double)


class OGrid(C.Structure):
    _fields_ = [("ndim", C.c_int32), ("n", C.c_int32 * 3), ("S", C.c_int32), ("uniform", C.c_int32),
                ("xv", c_double_p * 3), ("xc", c_double_p * 3), ("xvel", (c_double_p * 3) * 3),
                ("nvel", (C.c_int32 * 3) * 3)]


def build(force: bool = False) -> Path:
    src = HERE / "justpic_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        res = subprocess.run(["make", "-C", str(HERE), "-B"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.jpo_lerp.restype = C.c_double
        _lib.jpo_lerp.argtypes = [C.c_int, c_double_p, c_double_p]
        _lib.jpo_max_threads.restype = C.c_int
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _pp(arrs):
    out = (c_double_p * max(len(arrs), 1))()
    for i, a in enumerate(arrs):
        assert a.dtype == np.float64 and a.flags.c_contiguous
        out[i] = _dp(a)
    return out


class Oracle:
    """Grid + call wrappers.  xvi/xci: tuples of 1-D arrays; xi_vel[comp][dim]."""

    def __init__(self, xvi, xci, xi_vel, S, uniform):
        self.N = len(xvi)
        self.n = tuple(len(x) for x in xci)
        self.S = int(S)
        self._keep = []
        g = OGrid()
        g.ndim, g.S, g.uniform = self.N, self.S, 1 if uniform else 0
        for d in range(3):
            g.n[d] = self.n[d] if d < self.N else 1

        def ptr(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._keep.append(a)
            return _dp(a)

        for d in range(self.N):
            g.xv[d] = ptr(xvi[d])
            g.xc[d] = ptr(xci[d])
            for c in range(self.N):
                g.xvel[c][d] = ptr(xi_vel[c][d])
                g.nvel[c][d] = len(xi_vel[c][d])
        self.g = g
        self.C = int(np.prod(self.n))

    @staticmethod
    def set_threads(n):
        lib().jpo_set_threads(int(n))

    @staticmethod
    def max_threads():
        return int(lib().jpo_max_threads())

    def cell_shape(self, K=None):
        return ((self.S if K is None else K), *reversed(self.n))

    def init_particles(self, nxcell, seed):
        coords = [np.empty(self.cell_shape()) for _ in range(self.N)]
        index = np.empty(self.cell_shape(), dtype=np.uint8)
        rc = lib().jpo_init_particles(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), int(nxcell),
                                      C.c_uint64(int(seed)))
        assert rc == 0
        return coords, index

    def advect(self, coords, index, scheme, alpha, V, dt):
        return lib().jpo_advect(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), int(scheme),
                                C.c_double(alpha), _pp(V), C.c_double(dt))

    def advect_interp(self, coords, index, scheme, alpha, V, dt, interp):
        """interp: 1 = advection_LinP!, 2 = advection_MQS!"""
        return lib().jpo_advect_interp(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), int(scheme),
                                       C.c_double(alpha), _pp(V), C.c_double(dt), int(interp))

    def interp_velocity(self, V, p, cell1, interp):
        out = np.zeros(self.N)
        pp = np.ascontiguousarray(p, dtype=np.float64)
        cc = (C.c_int * 3)(*([int(c) for c in cell1] + [1] * (3 - len(cell1))))
        lib().jpo_interp_velocity(C.byref(self.g), _pp(V), _dp(pp), cc, int(interp), _dp(out))
        return out

    @staticmethod
    def set_move_policy(policy):
        """False / "reference": the reference's carried-over free-slot cursor; True / "compact" and "dense": the library's two
        optional policies (JP_OPT_MOVE_POLICY in include/justpic_c.h)."""
        lib().jpo_set_move_policy({False: 0, True: 1, "reference": 0, "compact": 1, "dense": 2}[policy])

    def move(self, coords, index, args):
        st = (C.c_int64 * 3)()
        rc = lib().jpo_move(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), _pp(args), len(args), st)
        assert rc == 0
        return tuple(int(v) for v in st)

    def clean(self, coords, index, args):
        return lib().jpo_clean(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), _pp(args), len(args))

    def force_injection(self, coords, index, pnew, fields, values):
        vals = (C.c_double * max(len(values), 1))(*[float(v) for v in values])
        return lib().jpo_force_injection(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), _pp(pnew), _pp(fields), vals,
                                         len(fields))

    def inject(self, coords, index, args, min_xcell, seed, step):
        n = C.c_int64()
        rc = lib().jpo_inject(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), _pp(args), len(args),
                              int(min_xcell), C.c_uint64(int(seed)), C.c_uint32(int(step)), C.byref(n))
        assert rc == 0
        return int(n.value)

    def inject_phase(self, coords, index, phases, args, fields, fkind, min_xcell, seed, step):
        n = C.c_int64()
        kinds = (C.c_int * max(len(args), 1))(*[int(k) for k in fkind])
        rc = lib().jpo_inject_phase(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), _dp(phases), _pp(args),
                                    _pp(fields), kinds, len(args), int(min_xcell), C.c_uint64(int(seed)), C.c_uint32(int(step)),
                                    C.byref(n))
        assert rc == 0
        return int(n.value)

    def grid2particle(self, coords, index, Fp, F):
        return lib().jpo_grid2particle(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), _dp(Fp), _dp(F))

    def grid2particle_flip(self, coords, index, Fp, F, F0, alpha):
        return lib().jpo_grid2particle_flip(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), _dp(Fp), _dp(F), _dp(F0),
                                            C.c_double(alpha))

    def subgrid_diffusion(self, coords, index, pT, T_grid, dT_grid, pT0, pdT, dt0, dT_subgrid, dt, d=1.0, centroid=False):
        """subgrid_diffusion! / subgrid_diffusion_centroid! (src/Physics/subgrid_diffusion.jl:55-113), composed
        from the restated kernels exactly as the reference composes them."""
        pT0[...] = pT                                                                       # memcopy_cellarray_kernel!
        if centroid:
            self.centroid2particle(coords, pT, T_grid)
        else:
            self.grid2particle(coords, index, pT, T_grid)
        lib().jpo_subgrid_kernel(C.byref(self.g), index.ctypes.data_as(C.c_void_p), _dp(pT), _dp(pT0), _dp(pdT), _dp(dt0),
                                 C.c_double(d), C.c_double(dt))
        if centroid:
            self.particle2centroid(coords, dT_subgrid, pdT)
        else:
            self.particle2grid(coords, index, dT_subgrid, pdT)
        nsub = (C.c_int * 3)(*(list(reversed(dT_subgrid.shape)) + [1] * (3 - dT_subgrid.ndim)))
        ndt = (C.c_int * 3)(*(list(reversed(dT_grid.shape)) + [1] * (3 - dT_grid.ndim)))
        lib().jpo_update_dT_subgrid(self.N, nsub, ndt, _dp(dT_subgrid), _dp(dT_grid))         # update_ΔT_subgrid_kernel!
        if centroid:
            self.centroid2particle(coords, pdT, dT_subgrid)
        else:
            self.grid2particle(coords, index, pdT, dT_subgrid)
        pT[...] = pT0 + pdT                                                                 # update_particle_temperature_kernel!

    def centroid2particle(self, coords, Fp, Fc):
        return lib().jpo_centroid2particle(C.byref(self.g), _pp(coords), _dp(Fp), _dp(Fc))

    def particle2grid(self, coords, index, F, Fp):
        return lib().jpo_particle2grid(C.byref(self.g), _pp(coords), index.ctypes.data_as(C.c_void_p), _dp(F), _dp(Fp))

    def particle2centroid(self, coords, Fc, Fp):
        return lib().jpo_particle2centroid(C.byref(self.g), _pp(coords), _dp(Fc), _dp(Fp))

    def phase_ratios_vertex(self, coords, ratios, phases, K):
        return lib().jpo_phase_ratios_vertex(C.byref(self.g), _pp(coords), _dp(ratios), _dp(phases), int(K))

    def phase_ratios_face(self, coords, ratios, phases, K, dim):
        return lib().jpo_phase_ratios_face(C.byref(self.g), _pp(coords), _dp(ratios), _dp(phases), int(K), int(dim))

    def phase_ratios_midpoint(self, coords, ratios, phases, K, plane):
        return lib().jpo_phase_ratios_midpoint(C.byref(self.g), _pp(coords), _dp(ratios), _dp(phases), int(K), int(plane))

    def phase_ratios_center(self, coords, ratios, phases, K):
        return lib().jpo_phase_ratios_center(C.byref(self.g), _pp(coords), _dp(ratios), _dp(phases), int(K))


def lerp(v, t):
    v = np.ascontiguousarray(v, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    return float(lib().jpo_lerp(len(t), _dp(v), _dp(t)))


def first_stage(scheme, alpha, dt, v, p):
    v = np.ascontiguousarray(v, dtype=np.float64); p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty_like(p)
    lib().jpo_first_stage(int(scheme), C.c_double(alpha), C.c_double(dt), len(p), _dp(v), _dp(p), _dp(out))
    return out


def second_stage(alpha, dt, v0, v1, p):
    v0 = np.ascontiguousarray(v0, dtype=np.float64); v1 = np.ascontiguousarray(v1, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty_like(p)
    lib().jpo_second_stage(C.c_double(alpha), C.c_double(dt), len(p), _dp(v0), _dp(v1), _dp(p), _dp(out))
    return out


def rand3(seed, purpose, step, cell, slot):
    r = np.empty(3)
    lib().jpo_rand3(C.c_uint64(int(seed)), C.c_uint32(purpose), C.c_uint32(step), C.c_uint32(cell), C.c_uint32(slot), _dp(r))
    return r


# ---- Array(CellArray) / CuArray(CellArray) (src/CellArrays/conversion.jl:32-43,
# ext/JustPICCUDAExt.jl:166-179): the device CellArray stores data[C, S, 1] (column-major: element
# (cell c, component s) at c + s*C, ext/JustPICCUDAExt.jl:26-30), the host CellArray data[1, S, C]
# (launch.jl:81); the conversion is permutedims(data, (3, 2, 1)) followed by copyto!, with
# convert(T, .) per element for the typed forms.  In this package's array convention
# (component axis first, then [nz,] ny, nx in C order) that is: move axis 0 to the end.
def cellarray_to_host_layout(a: np.ndarray, T=None) -> np.ndarray:
    out = np.ascontiguousarray(np.moveaxis(np.asarray(a), 0, -1))
    if a.dtype in (np.uint8, np.bool_):
        return out.astype(np.bool_)
    return out.astype(T) if T is not None else out


def cellarray_to_device_layout(a: np.ndarray, T=None) -> np.ndarray:
    out = np.ascontiguousarray(np.moveaxis(np.asarray(a), -1, 0))
    if a.dtype in (np.uint8, np.bool_):
        return out.astype(np.uint8)
    return out.astype(T) if T is not None else out
