"""Kernel-logic parity without a GPU, part 2: the cell-local kernels.  The loops of k_g2p / k_c2p / k_p2g (exact mode) / k_p2c /
k_phase / k_clean / k_advect_hi (justpic/jl_b200/csrc/justpic_sm100a.cu) are re-stated around the product's own per-particle
functions (jp_core.h compiled for the host, tests/emul/jp_emul.cpp) and must reproduce the oracle bit for bit -- grid2particle!,
centroid2particle!, particle2grid!, particle2centroid!, phase_ratios_center!, clean_particles!, advection_LinP! / advection_MQS!."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.emul.emul import Emul
from tests.problems import centre_field_linear, cfl_dt, make_grids, stream_velocity, vertex_field_linear

CASES = [(2, (13, 9), True), (2, (11, 14), False), (3, (7, 5, 6), True), (3, (5, 6, 4), False)]
ids = lambda c: f"{c[0]}D-{c[1]}-{'range' if c[2] else 'vector'}"


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def _state(case, S=20, nxcell=12, steps=3):
    """A few coupled oracle steps so that slots are ragged (holes, NaN slots, particles near faces)."""
    ndim, n, uniform = case
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.4)
    o = Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    e = Emul(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    co, idx = o.init_particles(nxcell, 5)
    V = stream_velocity(gr); dt = cfl_dt(gr, V, 0.8)
    f = np.where(idx > 0, co[0] * 2.0 - co[-1], np.nan)
    for it in range(steps):
        o.advect(co, idx, 1, 0.5, V, dt); o.move(co, idx, [f])
    return gr, o, e, co, idx, f, V, dt


@pytest.mark.parametrize("case", CASES, ids=ids)
def test_interpolation_kernels_match_oracle(case):
    gr, o, e, co, idx, f, V, dt = _state(case)
    T = vertex_field_linear(gr) + 0.25 * np.sin(7 * vertex_field_linear(gr, 0))
    Tc = centre_field_linear(gr) ** 2 + centre_field_linear(gr, 0)
    a, b = np.zeros_like(co[0]), np.zeros_like(co[0])
    o.grid2particle(co, idx, a, T); e.grid2particle(co, idx, b, T)
    assert same(a, b), "grid2particle"
    o.centroid2particle(co, a, Tc); e.centroid2particle(co, b, Tc)
    assert same(a, b), "centroid2particle"
    Fa, Fb = np.empty_like(T), np.empty_like(T)
    o.particle2grid(co, idx, Fa, a); e.particle2grid(co, idx, Fb, b)
    assert same(Fa, Fb), "particle2grid (exact summation order)"
    Ca, Cb = np.empty_like(Tc), np.empty_like(Tc)
    o.particle2centroid(co, Ca, a); e.particle2centroid(co, Cb, b)
    assert same(Ca, Cb), "particle2centroid"
    for K in (2, 5):
        ph = np.where(idx > 0, 1.0 + (np.floor(np.abs(co[0]) * 37) % K), 0.0)
        ra, rb = np.zeros(o.cell_shape(K)), np.zeros(o.cell_shape(K))
        o.phase_ratios_center(co, ra, ph, K); e.phase_ratios_center(co, rb, ph, K)
        assert same(ra, rb), f"phase_ratios_center K={K}"
        np.testing.assert_allclose(ra.sum(axis=0), 1.0, rtol=1e-13)


@pytest.mark.parametrize("case", CASES, ids=ids)
def test_clean_matches_oracle(case):
    gr, o, e, co, idx, f, V, dt = _state(case)
    o.advect(co, idx, 0, 0.0, V, dt)                       # leave particles outside their cells, no move
    A = [[c.copy() for c in co], idx.copy(), [f.copy()]]
    B = [[c.copy() for c in co], idx.copy(), [f.copy()]]
    o.clean(A[0], A[1], A[2]); e.clean(B[0], B[1], B[2])
    assert all(same(x, y) for x, y in zip(A[0], B[0])) and np.array_equal(A[1], B[1]) and same(A[2][0], B[2][0])
    assert int(A[1].sum()) < int(idx.sum())                # something was removed


@pytest.mark.parametrize("interp", [1, 2], ids=["LinP", "MQS"])
@pytest.mark.parametrize("case", CASES, ids=ids)
def test_advection_linp_mqs_match_oracle(case, interp):
    gr, o, e, co, idx, f, V, dt = _state(case, steps=1)
    for scheme, alpha in [(1, 0.5), (2, 0.0), (0, 0.0), (1, 2 / 3)]:
        A = [c.copy() for c in co]; B = [c.copy() for c in co]
        assert o.advect_interp(A, idx, scheme, alpha, V, 0.5 * dt, interp) == 0
        e.advect_interp(B, idx, scheme, alpha, V, 0.5 * dt, interp)
        assert all(same(x, y) for x, y in zip(A, B)), f"scheme {scheme} alpha {alpha}"
        co = A
