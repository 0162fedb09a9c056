"""phase_ratios_vertex! / phase_ratios_face! / phase_ratios_midpoint!.

CPU part: the C oracle against an independent, line-by-line numpy/Python transcription of the
Julia kernels (src/PhaseRatios/vertices.jl:15-107, midpoints.jl:26-82, :128-242) on small grids,
plus the reference's own property (ratios sum to 1, test/test_CellArrays.jl:109-112,176-183).
GPU part: the CUDA kernels against the oracle, bit for bit, through the public API."""
import itertools
import math

import numpy as np
import pytest

from oracle import oracle as O
from tests.problems import cfl_dt, make_grids, stream_velocity

OFF_FACE = {2: {"x": (1, 0), "y": (0, 1)}, 3: {"x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1)}}
OFF_MID = {"xy": (1, 1, 0), "yz": (0, 1, 1), "xz": (1, 0, 1)}


# ----------------------------------------------------------------------------- Julia transcription
def _dxi(gr, o, cell):
    """@dxi(di.vertex, I...): scalar x[2]-x[1] for range grids, diff(x)[I] for vector grids."""
    return [float(gr.xvi[d][1] - gr.xvi[d][0]) if gr.uniform else float(gr.xvi[d][cell[d] + 1] - gr.xvi[d][cell[d]])
            for d in range(gr.ndim)]


def _fma(a, b, c):
    """Correctly rounded a*b + c (math.fma needs Python >= 3.13)."""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _bilinear_weight(a, b, di):
    val = 1.0
    for x, y, d in zip(a, b, di):
        val *= _fma(-abs(x - y), 1.0 / d, 1.0)
    return val


def _cell_particles(co, ph, cell):
    idx = tuple(reversed(cell))
    for s in range(co[0].shape[0]):
        yield tuple(float(c[(s, *idx)]) for c in co), float(ph[(s, *idx)])


def _acc(w, x, phase):
    return [wk + (x if phase == k + 1 else math.copysign(0.0, x)) for k, wk in enumerate(w)]


def _norm(w, zero_nan):
    s = w[0]
    for v in w[1:]:
        s = s + v
    inv = math.inf if s == 0 else 1.0 / s
    out = []
    for v in w:
        r = v * inv if not (v == 0 and math.isinf(inv)) else math.nan
        out.append(0.0 if (zero_nan and math.isnan(r)) else r)
    return out


def julia_vertex(gr, co, ph, K):
    N, n = gr.ndim, gr.n
    out = np.zeros((K, *reversed([v + 1 for v in n])))
    for I in itertools.product(*[range(v + 1) for v in n]):
        xv = [float(gr.xvi[d][I[d]]) for d in range(N)]
        w = [0.0] * K
        for offs in itertools.product((-1, 0), repeat=N):          # offset_i outermost
            cell = tuple(I[d] + offs[d] for d in range(N))
            if any(c < 0 or c >= n[d] for d, c in enumerate(cell)):
                continue
            di = _dxi(gr, None, cell)
            for p, phase in _cell_particles(co, ph, cell):
                if any(math.isnan(v) for v in p):
                    continue
                if any(abs(p[d] - xv[d]) >= di[d] / 2 for d in range(N)):
                    continue
                w = _acc(w, _bilinear_weight(xv, p, di), phase)
        out[(slice(None), *reversed(I))] = _norm(w, False)
    return out


def julia_face(gr, co, ph, K, dim):
    N, n = gr.ndim, gr.n
    off = OFF_FACE[N][dim]
    out = np.zeros((K, *reversed([v + o for v, o in zip(n, off)])))
    for I in itertools.product(*[range(v) for v in n]):
        di = _dxi(gr, None, I)
        cen = [float(gr.xci[d][I[d]]) for d in range(N)]
        face = [cen[d] + di[d] * off[d] / 2 for d in range(N)]
        w = [0.0] * K
        for o2 in ((0,) * N, off):
            cell = tuple(min(I[d] + o2[d], n[d] - 1) for d in range(N))
            di = _dxi(gr, None, cell)
            for p, phase in _cell_particles(co, ph, cell):
                if any(math.isnan(v) for v in p):
                    continue
                if not all(abs(p[d] - face[d]) <= di[d] / 2 for d in range(N)):
                    continue
                w = _acc(w, _bilinear_weight(face, p, di), phase)
        out[(slice(None), *reversed([I[d] + off[d] for d in range(N)]))] = _norm(w, True)
        if any(off[d] * (I[d] + 1) == 1 for d in range(N)):
            face = [cen[d] - di[d] * off[d] / 2 for d in range(N)]
            w = [0.0] * K
            for p, phase in _cell_particles(co, ph, I):
                if any(math.isnan(v) for v in p):
                    continue
                if not all(abs(p[d] - face[d]) <= di[d] / 2 for d in range(N)):
                    continue
                w = _acc(w, _bilinear_weight(face, p, di), phase)
            out[(slice(None), *reversed(I))] = _norm(w, True)
    return out


def julia_midpoint(gr, co, ph, K, plane):
    n = gr.n
    off = OFF_MID[plane]
    MASK = ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1))
    out = np.zeros((K, *reversed([v + o for v, o in zip(n, off)])))

    def accumulate(I, mid):
        w = [0.0] * K
        for m in MASK:
            cell = tuple(min(I[d] + off[d] * m[d], n[d] - 1) for d in range(3))
            di = _dxi(gr, None, cell)
            for p, phase in _cell_particles(co, ph, cell):
                if any(math.isnan(v) for v in p):
                    continue
                if not all(abs(p[d] - mid[d]) <= di[d] / 2 for d in range(3)):
                    continue
                w = _acc(w, _bilinear_weight(mid, p, di), phase)
        return w

    for I in itertools.product(*[range(v) for v in n]):
        di = _dxi(gr, None, I)
        cen = [float(gr.xci[d][I[d]]) for d in range(3)]
        mid = [cen[d] + di[d] * off[d] / 2 for d in range(3)]
        out[(slice(None), *reversed([I[d] + off[d] for d in range(3)]))] = _norm(accumulate(I, mid), True)
        if any(off[d] * (I[d] + 1) == 1 for d in range(3)):
            ob = tuple(int(n[d] == off[d] * (I[d] + 1)) for d in range(3))
            for obi in ((0, 0, 0), ob):
                flip = tuple(0 - v for v in ob)
                di = _dxi(gr, None, I)
                mid = [cen[d] - (di[d] * off[d] * flip[d]) / 2 for d in range(3)]
                out[(slice(None), *reversed([I[d] + obi[d] for d in range(3)]))] = _norm(accumulate(I, mid), True)
    return out


# ----------------------------------------------------------------------------- shared set-up
def _state(ndim, n, uniform, K, seed=5, steps=2, holes=True):
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.3)
    S = 16
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    co, idx = o.init_particles(8, seed)
    V = stream_velocity(gr)
    dt = cfl_dt(gr, V, 0.6)
    rng = np.random.default_rng(seed)
    ph = np.where(idx > 0, rng.integers(1, K + 1, size=idx.shape), 0).astype(np.float64)
    for _ in range(steps):
        o.advect(co, idx, 1, 0.5, V, dt)
        o.move(co, idx, [ph])
    if holes:                                      # empty a few cells: vertices / faces with no particle in range
        for c in (tuple([0] * ndim), tuple(v // 2 for v in gr.n)):
            sl = (slice(None), *reversed(c))
            idx[sl] = 0
            ph[sl] = np.nan
            for a in co:
                a[sl] = np.nan
    return gr, o, co, idx, ph


def _same(a, b, what):
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    ok = ~np.isnan(b)
    assert np.array_equal(a[ok], b[ok]), f"{what}: {int((a[ok] != b[ok]).sum())} entries differ"


CASES = [(2, (6, 5), True), (2, (5, 7), False), (3, (4, 3, 5), True), (3, (3, 4, 3), False)]
cid = lambda c: f"{c[0]}D-{'x'.join(map(str, c[1]))}-{'range' if c[2] else 'vector'}"


@pytest.mark.parametrize("case", CASES, ids=cid)
def test_oracle_vertex_matches_julia_transcription(case):
    K = 3
    gr, o, co, idx, ph = _state(*case, K)
    out = np.full((K, *reversed([v + 1 for v in gr.n])), -7.0)
    assert o.phase_ratios_vertex(co, out, ph, K) == 0
    _same(out, julia_vertex(gr, co, ph, K), "phase_ratios_vertex")
    s = out.sum(axis=0)
    np.testing.assert_allclose(s[~np.isnan(s)], 1.0, rtol=1e-14)          # test_CellArrays.jl:109-112


@pytest.mark.parametrize("case", CASES, ids=cid)
def test_oracle_face_matches_julia_transcription(case):
    K = 3
    gr, o, co, idx, ph = _state(*case, K)
    for di, dim in enumerate("xyz"[:gr.ndim]):
        out = np.full((K, *reversed([v + (1 if d == di else 0) for d, v in enumerate(gr.n)])), -7.0)
        assert o.phase_ratios_face(co, out, ph, K, di) == 0
        _same(out, julia_face(gr, co, ph, K, dim), f"phase_ratios_face :{dim}")
        s = out.sum(axis=0)
        assert np.all((np.abs(s - 1.0) < 1e-14) | (s == 0.0))              # NaN -> 0 rows sum to 0


@pytest.mark.parametrize("case", [c for c in CASES if c[0] == 3], ids=cid)
def test_oracle_midpoint_matches_julia_transcription(case):
    K = 2
    gr, o, co, idx, ph = _state(*case, K)
    for pl, plane in enumerate(("xy", "yz", "xz")):
        off = OFF_MID[plane]
        out = np.zeros((K, *reversed([v + f for v, f in zip(gr.n, off)])))
        assert o.phase_ratios_midpoint(co, out, ph, K, pl) == 0
        _same(out, julia_midpoint(gr, co, ph, K, plane), f"phase_ratios_midpoint :{plane}")


# ----------------------------------------------------------------------------- GPU parity
GPU_CASES = [(2, (24, 17), True), (2, (19, 33), False), (3, (10, 9, 12), True), (3, (9, 7, 12), False)]


def _close(a, b, what, rtol=1e-12):
    """Fused mode: same NaN pattern, within the stated 1e-12 of the oracle."""
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    ok = ~np.isnan(b)
    np.testing.assert_allclose(a[ok], b[ok], rtol=rtol, atol=rtol, err_msg=what)


@pytest.mark.gpu
@pytest.mark.parametrize("case", GPU_CASES, ids=cid)
@pytest.mark.parametrize("K", [2, 4, 5])
@pytest.mark.parametrize("mode", ["literal", "fused"])
def test_gpu_update_phase_ratios(case, K, mode):
    import torch
    import justpic.jl_b200 as J
    gr, o, co, idx, ph = _state(*case, K, steps=3)
    grids = gr.grid_vel if gr.uniform else gr.xi_vel
    p = J.init_particles(J.CUDABackend, 8, 16, 4, *grids, seed=5)
    for d in range(gr.ndim):
        p.coords[d].copy_(torch.from_numpy(co[d]))
    p.index.copy_(torch.from_numpy(idx))
    phd = torch.from_numpy(ph).cuda()
    pr = J.PhaseRatios(J.CUDABackend, K, gr.n)
    for f in (pr.vertex, pr.Vx, pr.Vy, pr.Vz, pr.xy, pr.yz, pr.xz):
        f.fill_(-7.0)
    J.update_phase_ratios(pr, p, phd, mode=mode)
    n = gr.n
    cmp = _same if (mode == "literal" or K > 4) else _close      # fused falls back to the literal kernels for K > 4
    ref = np.zeros((K, *reversed(n))); o.phase_ratios_center(co, ref, ph, K)
    _same(pr.center.cpu().numpy(), ref, "center")                # centre ratios are bit-exact in both modes
    ref = np.full((K, *reversed([v + 1 for v in n])), -7.0); o.phase_ratios_vertex(co, ref, ph, K)
    cmp(pr.vertex.cpu().numpy(), ref, "vertex")
    for di, (dim, f) in enumerate(zip("xyz"[:gr.ndim], (pr.Vx, pr.Vy, pr.Vz))):
        ref = np.full((K, *reversed([v + (1 if d == di else 0) for d, v in enumerate(n)])), -7.0)
        o.phase_ratios_face(co, ref, ph, K, di)
        cmp(f.cpu().numpy(), ref, f"face :{dim}")
    if gr.ndim == 3:
        for pl, (plane, f) in enumerate(zip(("xy", "yz", "xz"), (pr.xy, pr.yz, pr.xz))):
            ref = np.full((K, *reversed([v + q for v, q in zip(n, OFF_MID[plane])])), -7.0)
            o.phase_ratios_midpoint(co, ref, ph, K, pl)
            cmp(f.cpu().numpy(), ref, f"midpoint :{plane}")


@pytest.mark.gpu
def test_gpu_phase_ratio_argument_errors():
    import torch
    import justpic.jl_b200 as J
    gr = make_grids(8, 2, True)
    p = J.init_particles(J.CUDABackend, 8, 16, 4, *gr.grid_vel, seed=1)
    pr = J.PhaseRatios(J.CUDABackend, 2, gr.n)
    ph, = J.init_cell_arrays(p, 1)
    with pytest.raises(ValueError):
        J.phase_ratios_face(pr.Vx, p, ph, "z")           # 2-D: :z is not a valid dimension
    with pytest.raises(ValueError):
        J.phase_ratios_midpoint(pr.xy, p, ph, "xy")      # midpoints are 3-D only
