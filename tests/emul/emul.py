"""ctypes front-end of the CPU emulation of the CUDA kernels (tests/emul/jp_emul.cpp).
TEST ONLY: compiles the product's jp_core.h for the host."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from justpic.jl_b200 import _cabi

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIB = HERE / "libjp_emul.so"
c_double_p = C.POINTER(C.c_double)


def build(force=False):
    deps = [HERE / "jp_emul.cpp", ROOT / "justpic/jl_b200/csrc/jp_core.h", ROOT / "justpic/jl_b200/csrc/jp_host_grid.h"]
    if force or not LIB.exists() or any(d.stat().st_mtime > LIB.stat().st_mtime for d in deps):
        cmd = ["g++", "-B/usr/lib/gcc/x86_64-linux-gnu/13/", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
               "-fno-fast-math", "-mfma", "-o", str(LIB), str(HERE / "jp_emul.cpp")]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("emulation build failed:\n" + res.stdout + res.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.jpe_create.restype = C.c_void_p
        _lib.jpe_create.argtypes = [C.POINTER(_cabi.GridDesc)]
        _lib.jpe_destroy.argtypes = [C.c_void_p]
        _lib.jpe_fast.argtypes = [C.c_void_p]
        _lib.jpe_inject.restype = C.c_int64
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _pp(arrs):
    out = (c_double_p * max(len(arrs), 1))()
    for i, a in enumerate(arrs):
        assert a.dtype == np.float64 and a.flags.c_contiguous
        out[i] = _dp(a)
    return out


class Emul:
    def __init__(self, xvi, xci, xi_vel, S, uniform):
        self.N = len(xvi)
        self.n = tuple(len(x) for x in xci)
        self.S = int(S)
        self._keep = []
        g = _cabi.GridDesc()
        g.ndim, g.S, g.uniform = self.N, self.S, 1 if uniform else 0
        for d in range(3):
            g.n[d] = self.n[d] if d < self.N else 1

        def ptr(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._keep.append(a)
            return _dp(a)

        for d in range(self.N):
            g.xv[d] = ptr(xvi[d]); g.xc[d] = ptr(xci[d])
            for c in range(self.N):
                g.xvel[c][d] = ptr(xi_vel[c][d]); g.nvel[c][d] = len(xi_vel[c][d])
        self.h = lib().jpe_create(C.byref(g))
        assert self.h, "jpe_create failed"

    def __del__(self):
        try:
            lib().jpe_destroy(C.c_void_p(self.h))
        except Exception:
            pass

    @property
    def fast(self):
        return bool(lib().jpe_fast(C.c_void_p(self.h)))

    def vkind(self):
        out = (C.c_int * 9)()
        lib().jpe_vkind(C.c_void_p(self.h), out)
        return np.array(out[:]).reshape(3, 3)

    def init_particles(self, nxcell, seed):
        shape = (self.S, *reversed(self.n))
        coords = [np.empty(shape) for _ in range(self.N)]
        index = np.empty(shape, dtype=np.uint8)
        rc = lib().jpe_init(C.c_void_p(self.h), _pp(coords), index.ctypes.data_as(C.c_void_p), int(nxcell), C.c_uint64(int(seed)))
        assert rc == 0
        return coords, index

    def advect(self, coords, index, scheme, alpha, V, dt, force_literal=False):
        return lib().jpe_advect(C.c_void_p(self.h), _pp(coords), index.ctypes.data_as(C.c_void_p), int(scheme),
                                C.c_double(alpha), _pp(V), C.c_double(dt), int(force_literal))

    def move(self, coords, index, args):
        st = (C.c_int64 * 3)()
        lib().jpe_move(C.c_void_p(self.h), _pp(coords), index.ctypes.data_as(C.c_void_p), _pp(args), len(args), st)
        return tuple(int(v) for v in st)

    def inject(self, coords, index, args, min_xcell, seed, step):
        return int(lib().jpe_inject(C.c_void_p(self.h), _pp(coords), index.ctypes.data_as(C.c_void_p), _pp(args), len(args),
                                    int(min_xcell), C.c_uint64(int(seed)), C.c_uint32(int(step))))

    def classify(self, ci, p):
        """(fast pre-filter, exact classification, literal move_kernel! route) of one particle stored in cell ``ci``."""
        ci3 = (C.c_int * 3)(*(list(ci) + [0] * (3 - len(ci))))
        p3 = (C.c_double * 3)(*(list(p) + [0.0] * (3 - len(p))))
        out = (C.c_int * 3)()
        lib().jpe_classify.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_int)]
        lib().jpe_classify(C.c_void_p(self.h), ci3, p3, out)
        return int(out[0]), int(out[1]), int(out[2])

    # ---- cell-local kernels (loops of justpic_sm100a.cu around the per-particle functions of jp_core.h)
    def _h(self):
        return C.c_void_p(self.h)

    def grid2particle(self, coords, index, Fp, F):
        lib().jpe_grid2particle(self._h(), _pp(coords), index.ctypes.data_as(C.c_void_p), _dp(Fp), _dp(F))

    def centroid2particle(self, coords, Fp, Fc):
        lib().jpe_centroid2particle(self._h(), _pp(coords), _dp(Fp), _dp(Fc))

    def particle2grid(self, coords, index, F, Fp):
        lib().jpe_particle2grid(self._h(), _pp(coords), index.ctypes.data_as(C.c_void_p), _dp(F), _dp(Fp))

    def particle2centroid(self, coords, Fc, Fp):
        lib().jpe_particle2centroid(self._h(), _pp(coords), _dp(Fc), _dp(Fp))

    def phase_ratios_center(self, coords, ratios, phases, K):
        lib().jpe_phase_ratios_center(self._h(), _pp(coords), _dp(ratios), _dp(phases), int(K))

    def clean(self, coords, index, args):
        lib().jpe_clean(self._h(), _pp(coords), index.ctypes.data_as(C.c_void_p), _pp(args), len(args))

    def advect_interp(self, coords, index, scheme, alpha, V, dt, interp):
        lib().jpe_advect_interp(self._h(), _pp(coords), index.ctypes.data_as(C.c_void_p), int(scheme), C.c_double(alpha), _pp(V),
                                C.c_double(dt), int(interp))
