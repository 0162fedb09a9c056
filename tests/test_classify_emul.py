"""move_particles! classification without a GPU: jp_classify_fast / jp_classify_particle (justpic/jl_b200/csrc/jp_core.h,
the per-particle code of k_move_classify3 and of the advection -> move hand-off) compiled for the host and checked against the
literal route of move_kernel! (src/Particles/move_safe.jl:86-106: strict isincell against fl(x + dx), strict indomain, seeded
bisection) on adversarial positions: on vertices, one ulp either side, inside the fl(x + dx) gap, at the pre-filter's 1e-4
margin, in every neighbour, two cells away, outside the domain, NaN and Inf.

Contract: (1) whenever the single-precision pre-filter answers, it gives the exact routine's code; (2) whenever the exact routine
answers with a plannable code (stay / delete / one of the 3^N - 1 neighbours) the literal route gives the same; (3) whenever it
hands over to the direct sweeps (JP_CLS_CPLX + r), the literal route is indeed one of the cases the planner cannot express."""
import itertools

import numpy as np
import pytest

from tests.emul.emul import Emul
from tests.problems import make_grids

STAY, DELETE, CPLX = 28, 27, 28


def _positions(x, i, rng):
    """candidate coordinates for a particle stored in cell i of vertex vector x"""
    n = len(x) - 1
    dx = x[i + 1] - x[i]
    out = [x[i] + dx * u for u in (0.5, 1e-4, 0.99995e-4, 1.00005e-4, 1 - 1e-4, 1 - 0.9999e-4, 1e-9, 1 - 1e-9)]
    for v in (x[i], x[i + 1], x[i] + dx):                     # the vertices and the computed upper edge fl(x + dx)
        out += [v, np.nextafter(v, -np.inf), np.nextafter(v, np.inf), np.nextafter(np.nextafter(v, np.inf), np.inf)]
    for k in (-2, -1, 1, 2, 3):                               # neighbours, far cells, beyond the domain
        j = i + k
        lo = x[j] if 0 <= j <= n else x[0] + j * dx
        out += [lo + 0.3 * dx, lo + 1e-5 * dx, lo]
    out += [x[0], x[-1], np.nextafter(x[0], np.inf), np.nextafter(x[-1], -np.inf), x[0] - 1e-3, x[-1] + 1e-3, np.nan, np.inf, -np.inf]
    out += list(x[i] + dx * rng.uniform(-1.2, 2.2, 6))
    return out


@pytest.mark.parametrize("ndim,n,uniform,exact", [(2, (7, 5), True, False), (2, (8, 8), True, True), (3, (5, 4, 6), True, False),
                                                 (2, (7, 5), False, False), (3, (4, 5, 3), False, False)])
def test_classification_three_ways(ndim, n, uniform, exact):
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.4, exact=exact)
    e = Emul(gr.xvi, gr.xci, gr.xi_vel, 8, uniform)
    rng = np.random.default_rng(5)
    xs = [np.asarray(x, dtype=np.float64) for x in gr.xvi]
    cells = [(0,) * ndim, tuple(k - 1 for k in n), tuple(k // 2 for k in n), tuple(min(1, k - 1) for k in n)]
    n_fast = n_plan = n_cplx = 0
    for ci in cells:
        cand = [_positions(xs[d], ci[d], rng) for d in range(ndim)]
        # every adversarial value in one dimension against a few values in the others
        others = [[c[0], c[2], c[9], c[17], c[-4]] for c in cand]
        for d in range(ndim):
            pools = [cand[k] if k == d else others[k] for k in range(ndim)]
            for p in itertools.product(*pools):
                fast, exact_code, lit = e.classify(ci, p)
                if fast >= 0:
                    n_fast += 1
                    assert fast == exact_code, (ci, p, fast, exact_code)
                if exact_code <= STAY:                      # stay / delete / a neighbour code
                    n_plan += 1
                    assert exact_code == lit, (ci, p, exact_code, lit)
                else:
                    n_cplx += 1
                    assert exact_code in (CPLX + 1, CPLX + 2, CPLX + 3)
                    assert lit > STAY or lit == exact_code or _on_vertex(p, xs), (ci, p, exact_code, lit)
    assert n_plan > 300 and n_cplx > 50
    assert (n_fast > 100) if uniform else n_fast == 0       # the pre-filter only runs on range grids


def _on_vertex(p, xs):
    """A particle exactly on a grid vertex: the exact routine refuses it (strict comparisons on both sides), while the literal
    bisection resolves the tie to one side -- one of the cases that must go to the direct sweeps."""
    return any(np.isfinite(v) and (np.asarray(x) == v).any() for v, x in zip(p, xs))


@pytest.mark.parametrize("ndim,n,uniform", [(2, (16, 9), True), (3, (6, 5, 7), True), (3, (6, 5, 7), False)])
def test_classification_random_sweep(ndim, n, uniform):
    """Random storage cells and positions (anywhere from one cell outside the domain to the other side), plus positions a few
    ulps / a fraction of the pre-filter margin from random vertices: the same three-way contract."""
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.35)
    e = Emul(gr.xvi, gr.xci, gr.xi_vel, 8, uniform)
    xs = [np.asarray(x, dtype=np.float64) for x in gr.xvi]
    rng = np.random.default_rng(11)
    agree_fast = plannable = 0
    for it in range(30000):
        ci = tuple(int(rng.integers(0, n[d])) for d in range(ndim))
        p = []
        for d in range(ndim):
            x = xs[d]; dx = x[ci[d] + 1] - x[ci[d]]
            mode = rng.integers(0, 4)
            if mode == 0:   v = x[ci[d]] + dx * rng.uniform(-1.5, 2.5)                     # own cell and its neighbourhood
            elif mode == 1: v = rng.uniform(x[0] - dx, x[-1] + dx)                          # anywhere
            elif mode == 2:                                                                  # a few ulps from a nearby vertex
                v = x[int(np.clip(ci[d] + rng.integers(-1, 3), 0, n[d]))]
                for _ in range(int(rng.integers(0, 4))): v = np.nextafter(v, rng.choice([-np.inf, np.inf]))
            else:           v = x[int(np.clip(ci[d] + rng.integers(0, 2), 0, n[d]))] + dx * 1e-4 * rng.uniform(-1.5, 1.5)
            p.append(float(v))
        fast, exact_code, lit = e.classify(ci, p)
        if fast >= 0:
            agree_fast += 1
            assert fast == exact_code, (ci, p, fast, exact_code)
        if exact_code <= STAY:
            plannable += 1
            assert exact_code == lit, (ci, p, exact_code, lit)
    assert plannable > 10000 and ((agree_fast > 500) if uniform else agree_fast == 0)
