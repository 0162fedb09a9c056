"""advection_LinP! / advection_MQS!.
CPU: the oracle's LinP / MQS velocity reconstruction against an independent line-by-line Python
transcription of the Julia sources (src/Interpolations/MQS.jl, src/Particles/Advection/
advection_LinP.jl:96-391, advection_MQS.jl:96-124) at random points; the reference's own property
(test/test_2D.jl:191-197: the result does not depend on the seed cell).
GPU: advection_LinP / advection_MQS against the oracle, bit for bit."""
import itertools
from fractions import Fraction

import numpy as np
import pytest

from oracle import oracle as O
from tests.problems import cfl_dt, make_grids, stream_velocity


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def lerp1(t, v0, v1):
    return fma(t, v1, fma(-t, v0, v0))


def lerp(v, t):
    v = list(v)
    for td in t:
        v = [lerp1(td, v[2 * q], v[2 * q + 1]) for q in range(len(v) // 2)]
    return v[0]


def bisect(px, x, seed):                      # find_parent_cell_bisection, 1-based
    lo, hi = 1, len(x)
    while True:
        if x[seed - 1] <= px <= x[seed]:
            return seed
        if x[seed - 1] < px:
            lo, seed = seed, (hi + seed) // 2
        else:
            hi, seed = seed, (lo + seed) // 2


class JuliaF:
    """1-based view of a ([nz,] ny, nx) numpy array as Julia's F[i, j[, k]]."""

    def __init__(self, a):
        self.a = a
        self.size = tuple(reversed(a.shape))

    def __getitem__(self, ijk):
        return float(self.a[tuple(reversed([q - 1 for q in ijk]))])


def corners(F, idx):
    N = len(idx)
    return [F[tuple(idx[d] + ((q >> d) & 1) for d in range(N))] for q in range(2 ** N)]


def mqs4(F, v, t, ijk, kind):
    """The 4-tuple MQS methods (2-D x/y; 3-D x/y/z), MQS.jl:1-60, :80-158."""
    t1, t2 = t
    half = 0.5
    N = len(ijk)

    def at(di, dj, dk=0):
        q = [ijk[0] + di, ijk[1] + dj] + ([ijk[2] + dk] if N == 3 else [])
        return F[tuple(q)]

    def corr(tq, v0, v1, v2):
        return (half * (tq - half) ** 2) * (fma(-2.0, v1, v0) + v2)

    if kind in ("x", "z"):
        lerp_bot, lerp_top = lerp(v[0:2], (t1,)), lerp(v[2:4], (t1,))
        top = (0, 1, 0) if kind == "x" else (0, 0, 1)
        v0, v1, v2 = (at(-1, 0, 0), v[0], v[1]) if t1 < half else (v[0], v[1], at(2, 0, 0))
        cb = corr(t1, v0, v1, v2)
        v0, v1, v2 = (at(-1, top[1], top[2]), v[2], v[3]) if t1 < half else (v[2], v[3], at(2, top[1], top[2]))
        ct = corr(t1, v0, v1, v2)
        return lerp((lerp_bot + cb, lerp_top + ct), (t2,))
    vl, vr = (v[0], v[2]), (v[1], v[3])
    ll, lr = lerp(vl, (t2,)), lerp(vr, (t2,))
    v0, v1, v2 = (at(0, -1), *vl) if t2 < half else (*vl, at(0, 2))
    cl = corr(t2, v0, v1, v2)
    v0, v1, v2 = (at(1, -1), *vr) if t2 < half else (*vr, at(1, 2))
    cr = corr(t2, v0, v1, v2)
    return lerp((ll + cl, lr + cr), (t1,))


def mqs(F, v, t, idx, comp):
    N = len(idx)
    if N == 2:
        return mqs4(F, v, t, idx, "xy"[comp])
    if comp in (0, 1):
        bot = mqs4(F, v[0:4], t[0:2], idx, "xy"[comp])
        top = mqs4(F, v[4:8], t[0:2], idx, "xy"[comp])
        return lerp((bot, top), (t[2],))
    front = mqs4(F, (v[0], v[1], v[4], v[5]), (t[0], t[2]), idx, "z")
    back = mqs4(F, (v[2], v[3], v[6], v[7]), (t[0], t[2]), idx, "z")
    return lerp((front, back), (t[1],))


AUG = {0: (((-1, 0, 1),) * 4, ((0, 0, 0), (1, 1, 1), (0, 0, 0), (1, 1, 1)), ((0, 0, 0), (0, 0, 0), (1, 1, 1), (1, 1, 1))),
       1: (((0, 0, 0), (1, 1, 1), (0, 0, 0), (1, 1, 1)), ((-1, 0, 1),) * 4, ((0, 0, 0), (0, 0, 0), (1, 1, 1), (1, 1, 1))),
       2: (((0, 0, 0), (0, 0, 0), (1, 1, 1), (1, 1, 1)), ((0, 0, 0), (1, 1, 1), (0, 0, 0), (1, 1, 1)), ((-1, 0, 1),) * 4)}


def interpolate_V_to_P(F, xc, p, dxi, comp, idx):
    N = len(idx)
    ijk = list(idx)
    ijk[comp] += int(p[comp] > xc[comp] + dxi[comp] / 2)
    oi, oj, ok = AUG[comp]
    clamp = lambda x, lo, hi: hi if x > hi else (lo if x < lo else x)
    rows = [(0, 0, 0), (1, 1, 1)] if N == 2 else [(0, 0, 0), (1, 1, 1), (2, 0, 2), (3, 1, 3)]
    av = []
    for (ri, rj, rk) in rows:
        f = []
        for m in range(3):
            q = [clamp(ijk[0] + oi[ri][m], 1, F.size[0]), clamp(ijk[1] + oj[rj][m], 1, F.size[1])]
            if N == 3:
                q.append(clamp(ijk[2] + ok[rk][m], 1, F.size[2]))
            f.append(F[tuple(q)])
        av += [(f[0] + f[1]) / 2, (f[2] + f[1]) / 2]
    if comp == 0:
        return av
    if N == 2:
        return [av[0], av[2], av[1], av[3]]
    return [av[0], av[2], av[1], av[3], av[4], av[6], av[5], av[7]]


def julia_velocity(gr, V, p, cell1, interp):
    N = gr.ndim
    out = []
    for c in range(N):
        grid = gr.xi_vel[c]
        if not all(grid[d][0] <= p[d] <= grid[d][-1] for d in range(N)):
            out.append(np.inf)
            continue
        idx = [bisect(p[d], grid[d], cell1[d]) for d in range(N)]
        xc = [float(grid[d][idx[d] - 1]) for d in range(N)]
        dxi = [float(grid[d][1] - grid[d][0]) if gr.uniform else float(grid[d][idx[d]] - grid[d][idx[d] - 1]) for d in range(N)]
        F = JuliaF(V[c])
        Fi = corners(F, idx)
        t = [(p[d] - xc[d]) * (1.0 / dxi[d]) for d in range(N)]
        VL = lerp(Fi, t)
        interior = all(1 < idx[d] < F.size[d] - 1 for d in range(N))
        if interp == 0 or not interior:
            out.append(VL)
        elif interp == 2:
            out.append(mqs(F, Fi, t, idx, c))
        else:
            FP = interpolate_V_to_P(F, xc, p, dxi, c, idx)
            xP = list(xc)
            off = 1 - 2 * int(p[c] < xc[c] + dxi[c] / 2)
            xP[c] = xc[c] + off * dxi[c] / 2
            tP = [(p[d] - xP[d]) * (1.0 / dxi[d]) for d in range(N)]
            VP = lerp(FP, tP)
            A = 2 / 3
            out.append(A * VL + (1 - A) * VP)
    return np.array(out)


def _problem(ndim, n, uniform):
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.3)
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, 8, uniform)
    rng = np.random.default_rng(4)
    V = [v + 0.3 * rng.standard_normal(v.shape) for v in stream_velocity(gr, amp=1.0)]      # rough field: every stencil node matters
    return gr, o, [np.ascontiguousarray(v) for v in V]


CASES = [(2, (7, 6), True), (2, (6, 8), False), (3, (6, 5, 7), True), (3, (5, 6, 5), False)]
cid = lambda c: f"{c[0]}D-{'x'.join(map(str, c[1]))}-{'range' if c[2] else 'vector'}"


@pytest.mark.parametrize("case", CASES, ids=cid)
@pytest.mark.parametrize("interp", [1, 2], ids=["LinP", "MQS"])
def test_oracle_interpolant_matches_julia_transcription(case, interp):
    gr, o, V = _problem(*case)
    rng = np.random.default_rng(7)
    N = gr.ndim
    for _ in range(300):
        cell = [int(rng.integers(0, gr.n[d])) for d in range(N)]
        # a point in the cell or up to one cell away (as in a second RK stage), sometimes outside the domain
        p = [float(gr.xvi[d][cell[d]] + (gr.xvi[d][cell[d] + 1] - gr.xvi[d][cell[d]]) * rng.uniform(-0.9, 1.9)) for d in range(N)]
        cell1 = [c + 1 for c in cell]
        got = o.interp_velocity(V, p, cell1, interp)
        want = julia_velocity(gr, V, p, cell1, interp)
        assert np.array_equal(got, want), (p, cell1, got, want)


@pytest.mark.parametrize("interp", [1, 2], ids=["LinP", "MQS"])
def test_oracle_interpolant_seed_independent(interp):
    # test/test_2D.jl:168-197 (refined grid, p = (0.22, 0.48), seed (3,3) vs the corrected seed)
    xv = np.array([0.0, 0.1, 0.3, 0.6, 1.0]); yv = np.linspace(0, 1, 5)
    xc, yc = 0.5 * (xv[1:] + xv[:-1]), 0.5 * (yv[1:] + yv[:-1])
    ext = lambda c: np.concatenate(([c[0] - (c[1] - c[0])], c, [c[-1] + (c[-1] - c[-2])]))
    xi_vel = ((xv, ext(yc)), (ext(xc), yv))
    o = O.Oracle((xv, yv), (xc, yc), xi_vel, 8, False)
    Vx = np.ascontiguousarray((2 * xi_vel[0][0][None, :] + xi_vel[0][1][:, None]))
    Vy = np.ascontiguousarray((xi_vel[1][0][None, :] - 3 * xi_vel[1][1][:, None]))
    p = (0.22, 0.48)
    a = o.interp_velocity([Vx, Vy], p, (3, 3), interp)
    corrected = (bisect(p[0], xv, 3), bisect(p[1], yv, 3))
    b = o.interp_velocity([Vx, Vy], p, corrected, interp)
    np.testing.assert_allclose(a, b, rtol=1.5e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(2, (24, 17), True), (2, (19, 33), False), (3, (10, 9, 12), True), (3, (9, 7, 12), False)], ids=cid)
@pytest.mark.parametrize("interp", ["LinP", "MQS"])
@pytest.mark.parametrize("method", ["euler", "rk2", "rk4"])
def test_gpu_advection_interpolants(case, interp, method):
    import torch
    import justpic.jl_b200 as J
    gr, o, V = _problem(*case)
    grids = gr.grid_vel if gr.uniform else gr.xi_vel
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, 16, gr.uniform)
    p = J.init_particles(J.CUDABackend, 8, 16, 4, *grids, seed=9)
    co, idx = o.init_particles(8, 9)
    Vd = [torch.from_numpy(v).cuda() for v in V]
    m = {"euler": (J.Euler(), 0, 0.0), "rk2": (J.RungeKutta2(2 / 3), 1, 2 / 3), "rk4": (J.RungeKutta4(), 2, 0.0)}[method]
    fn = J.advection_LinP if interp == "LinP" else J.advection_MQS
    code = 1 if interp == "LinP" else 2
    for cfl in (0.4, 1.3):
        dt = cfl_dt(gr, V, cfl)
        fn(p, m[0], Vd, dt)
        assert o.advect_interp(co, idx, m[1], m[2], V, dt, code) == 0
        for d in range(gr.ndim):
            a = p.coords[d].cpu().numpy()
            assert np.array_equal(a, co[d], equal_nan=True), f"{interp} {method} cfl {cfl}: coords[{d}] differ at {int((~((a == co[d]) | (np.isnan(a) & np.isnan(co[d])))).sum())}"
        J.move_particles(p); o.move(co, idx, [])
        assert np.array_equal(p.index.cpu().numpy(), idx)
