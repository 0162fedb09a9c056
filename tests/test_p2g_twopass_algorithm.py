"""particle2grid!, default mode, as an ALGORITHM on the CPU: per-cell partial sums towards the cell's 2^N corner nodes in slot
order (k_p2g_cell), then each node adds the partials of its <= 2^N cells in the reference's (k, j, i) order (k_p2g_node) -- the
same terms as the reference's single running sum (src/Interpolations/particle_to_grid.jl:37-151), associated differently, with
the weight evaluated as 1 / sum(d^2) instead of inv(sqrt(sum(d^2))^2).  Must agree with the oracle within the stated 1e-12
 -- measured here: a few 1e-16."""
import itertools

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.problems import cfl_dt, make_grids, stream_velocity


def twopass(gr, S, co, idx, Fp, fast_weight=True):
    N = gr.ndim
    n = list(gr.n)
    xv = [np.asarray(x, dtype=np.float64) for x in gr.xvi]
    NQ = 2 ** N
    shape_c = tuple(reversed(n))
    PW = np.zeros((NQ,) + shape_c); PWF = np.zeros((NQ,) + shape_c)
    for ci in itertools.product(*[range(k) for k in n]):                      # pass 1, thread = cell
        cidx = tuple(reversed(ci))
        for s in range(S):
            if not idx[(s,) + cidx]:
                continue
            p = [co[d][(s,) + cidx] for d in range(N)]
            f = Fp[(s,) + cidx]
            for q in range(NQ):
                ss = np.float64(0.0)
                for d in range(N):
                    a = xv[d][ci[d] + ((q >> d) & 1)] - p[d]
                    ss = ss + a * a if d else a * a
                w = 1.0 / ss if fast_weight else 1.0 / (np.sqrt(ss) * np.sqrt(ss))
                PW[(q,) + cidx] += w
                PWF[(q,) + cidx] = w * f + PWF[(q,) + cidx]
    F = np.empty(tuple(k + 1 for k in reversed(n)))
    for nd in itertools.product(*[range(k + 1) for k in n]):                 # pass 2, thread = node
        w = wF = np.float64(0.0)
        offs = list(itertools.product(*[(-1, 0)] * N))
        for o in sorted(offs, key=lambda t: tuple(reversed(t))):              # k outermost, i innermost
            c = [nd[d] + o[d] for d in range(N)]
            if any(c[d] < 0 or c[d] >= n[d] for d in range(N)):
                continue
            q = sum((1 if o[d] < 0 else 0) << d for d in range(N))
            w = w + PW[(q,) + tuple(reversed(c))]; wF = wF + PWF[(q,) + tuple(reversed(c))]
        with np.errstate(invalid="ignore", divide="ignore"):
            F[tuple(reversed(nd))] = wF / w
    return F


@pytest.mark.parametrize("ndim,n,uniform", [(2, (9, 7), True), (2, (8, 6), False), (3, (5, 4, 4), True), (3, (4, 5, 3), False)])
@pytest.mark.parametrize("fast_weight", [True, False])
def test_twopass_particle2grid_within_stated_tolerance(ndim, n, uniform, fast_weight):
    S = 16
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.4)
    o = Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    co, idx = o.init_particles(8, 4)
    V = stream_velocity(gr); dt = cfl_dt(gr, V, 0.8)
    Fp = np.where(idx > 0, np.sin(5 * co[0]) + 2.0 * co[-1], 0.0)
    for _ in range(2):
        o.advect(co, idx, 1, 0.5, V, dt); o.move(co, idx, [Fp])
    ref = np.empty(tuple(k + 1 for k in reversed(n)))
    o.particle2grid(co, idx, ref, Fp)
    got = twopass(gr, S, co, idx, Fp, fast_weight)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = np.isfinite(ref)
    scale = np.abs(ref[ok]).max()
    err = np.abs(got[ok] - ref[ok]).max() / scale
    assert err < 1e-12, err
    assert err < 1e-14                                   # in practice a few ulp
