"""Kernel-logic parity without a GPU: the CUDA kernels' per-thread code
(justpic/jl_b200/csrc/jp_core.h) compiled for the host and driven with the
kernels' decomposition (tests/emul/jp_emul.cpp) must reproduce the oracle
bit-for-bit -- including the fast velocity-interpolation path, the
classify + occupancy-word colour sweeps of move_particles! and the
flag + sweep structure of inject_particles!."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.emul.emul import Emul
from tests.problems import cfl_dt, make_grids, stream_velocity, vertex_field_linear


def eq(a, b):
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


def pair(gr, S):
    return (Oracle(gr.xvi, gr.xci, gr.xi_vel, S, gr.uniform), Emul(gr.xvi, gr.xci, gr.xi_vel, S, gr.uniform))


CASES = [
    # ndim, n, uniform, stretch, S, nxcell, min_xcell, cfl
    (2, 24, True, 0.0, 24, 12, 8, 0.75),
    (3, 10, True, 0.0, 24, 12, 8, 0.9),
    (2, 20, False, 0.5, 20, 12, 10, 0.9),
    (3, 9, False, 0.4, 20, 10, 8, 0.5),
    (2, 17, True, 0.0, 12, 12, 6, 0.95),     # tight storage: drops
    (3, (7, 5, 6), True, 0.0, 16, 8, 8, 2.5),  # > 1-cell moves: literal fallback everywhere
    (2, (5, 33), True, 0.0, 64, 40, 20, 0.6),  # S = 64 (full occupancy word)
    (2, (7, 9), True, 0.0, 80, 60, 50, 0.7),   # max_xcell > 64 (test/test_2D.jl:437): literal wide move / inject
    (3, (4, 5, 3), False, 0.3, 150, 125, 100, 0.6),  # test/test_3D.jl:369,424
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}D-n{c[1]}-{'rng' if c[2] else 'vec'}-S{c[4]}-cfl{c[7]}")
def test_trajectory_parity(case):
    ndim, n, uniform, stretch, S, nxc, minx, cfl = case
    gr = make_grids(n, ndim, uniform=uniform, stretch=stretch)
    o, e = pair(gr, S)
    assert e.fast
    co, idx = o.init_particles(nxc, 7)
    ce, ie = e.init_particles(nxc, 7)
    assert eq(co, ce) and np.array_equal(idx, ie)
    T = vertex_field_linear(gr)
    pT = np.zeros_like(co[0]); ph = np.zeros_like(co[0])
    o.grid2particle(co, idx, pT, T)
    ph[:] = np.where(idx > 0, 1.0 + (co[0] < co[-1]), 0.0)
    V = stream_velocity(gr)
    dt = cfl_dt(gr, V, cfl)
    A = [[a.copy() for a in co], idx.copy(), [pT.copy(), ph.copy()]]
    B = [[a.copy() for a in co], idx.copy(), [pT.copy(), ph.copy()]]
    schemes = [(1, 0.5), (2, 0.0), (1, 2 / 3), (0, 0.0)]
    total_moved = 0
    for it in range(10):
        sch = schemes[it % 4]
        o.advect(A[0], A[1], sch[0], sch[1], V, dt)
        e.advect(B[0], B[1], sch[0], sch[1], V, dt)
        assert eq(A[0], B[0]), f"advect diverged at step {it}"
        sa = o.move(A[0], A[1], A[2]); sb = e.move(B[0], B[1], B[2])
        assert sa == sb and eq(A[0], B[0]) and np.array_equal(A[1], B[1]) and eq(A[2], B[2]), f"move diverged at step {it}"
        total_moved += sa[0]
        ia = o.inject(A[0], A[1], A[2], minx, 7, it); ib = e.inject(B[0], B[1], B[2], minx, 7, it)
        assert ia == ib and eq(A[0], B[0]) and np.array_equal(A[1], B[1]) and eq(A[2], B[2]), f"inject diverged at step {it}"
    assert total_moved > 0


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("uniform", [True, False])
def test_fast_equals_literal_interpolation(ndim, uniform):
    gr = make_grids(9 if ndim == 3 else 21, ndim, uniform=uniform, stretch=0.45)
    o, e = pair(gr, 16)
    co, idx = o.init_particles(16, 11)
    V = stream_velocity(gr)
    for cfl in (0.3, 0.99, 1.7):
        dt = cfl_dt(gr, V, cfl)
        for scheme, alpha in [(0, 0.0), (1, 0.5), (1, 0.25), (2, 0.0)]:
            a = [x.copy() for x in co]; b = [x.copy() for x in co]; c = [x.copy() for x in co]
            o.advect(a, idx, scheme, alpha, V, dt)
            e.advect(b, idx, scheme, alpha, V, dt, force_literal=False)
            e.advect(c, idx, scheme, alpha, V, dt, force_literal=True)
            assert eq(a, b) and eq(a, c)


@pytest.mark.parametrize("ndim", [2, 3])
def test_ties_on_faces_vertices_and_centres(ndim):
    """Particles exactly on cell faces / vertices / centres / domain boundary,
    NaN and Inf coordinates: the fast path must fall back, move must re-slot or
    delete exactly as the reference's strict/inclusive comparisons dictate."""
    gr = make_grids(6, ndim, uniform=True)
    S = 12
    o, e = pair(gr, S)
    co, idx = o.init_particles(8, 5)
    rng = np.random.default_rng(1)
    xs = [np.concatenate([gr.xvi[d], gr.xci[d], [np.nan, np.inf, -np.inf, gr.xvi[d][0] - 1e-3, gr.xvi[d][-1] + 1e-3,
                                               np.nextafter(gr.xvi[d][2], 1.0), np.nextafter(gr.xvi[d][2], -1.0)]]) for d in range(ndim)]
    live = np.argwhere(idx > 0)
    pick = rng.choice(len(live), size=len(live) // 3, replace=False)
    for t in live[pick]:
        for d in range(ndim):
            if rng.random() < 0.6:
                # snap to a special value near the particle's own cell (so moves stay local) or anywhere
                cell = t[ndim - d]
                cand = [gr.xvi[d][cell], gr.xvi[d][cell + 1], gr.xci[d][cell]] if rng.random() < 0.8 else list(xs[d])
                co[d][tuple(t)] = cand[rng.integers(len(cand))]
    V = stream_velocity(gr)
    dt = cfl_dt(gr, V, 0.5)
    pT = np.where(idx > 0, rng.random(idx.shape), 0.0)
    A = [[a.copy() for a in co], idx.copy(), [pT.copy()]]
    B = [[a.copy() for a in co], idx.copy(), [pT.copy()]]
    # move first (ties in place), then advect from tie positions, then move/inject again
    sa = o.move(A[0], A[1], A[2]); sb = e.move(B[0], B[1], B[2])
    assert sa == sb and eq(A[0], B[0]) and np.array_equal(A[1], B[1]) and eq(A[2], B[2])
    assert sa[2] > 0            # some were outside the domain / NaN / Inf
    A = [[a.copy() for a in co], idx.copy(), [pT.copy()]]
    B = [[a.copy() for a in co], idx.copy(), [pT.copy()]]
    for scheme, alpha in [(1, 0.5), (2, 0.0), (0, 0.0)]:
        o.advect(A[0], A[1], scheme, alpha, V, dt); e.advect(B[0], B[1], scheme, alpha, V, dt)
        assert eq(A[0], B[0])
        sa = o.move(A[0], A[1], A[2]); sb = e.move(B[0], B[1], B[2])
        assert sa == sb and eq(A[0], B[0]) and np.array_equal(A[1], B[1]) and eq(A[2], B[2])
        ia = o.inject(A[0], A[1], A[2], 8, 5, scheme); ib = e.inject(B[0], B[1], B[2], 8, 5, scheme)
        assert ia == ib and eq(A[0], B[0]) and np.array_equal(A[1], B[1]) and eq(A[2], B[2])


def test_outflow_deletes_particles():
    gr = make_grids(8, 2, uniform=True)
    o, e = pair(gr, 16)
    co, idx = o.init_particles(8, 2)
    V = [np.full_like(v, 1.0) for v in stream_velocity(gr)]     # uniform flow towards +x,+y
    dt = 0.6 * (gr.xvi[0][1] - gr.xvi[0][0])
    A = [[a.copy() for a in co], idx.copy(), []]
    B = [[a.copy() for a in co], idx.copy(), []]
    deleted = 0
    for it in range(6):
        o.advect(A[0], A[1], 1, 0.5, V, dt); e.advect(B[0], B[1], 1, 0.5, V, dt)
        assert eq(A[0], B[0])
        sa = o.move(A[0], A[1], A[2]); sb = e.move(B[0], B[1], B[2])
        assert sa == sb and eq(A[0], B[0]) and np.array_equal(A[1], B[1])
        deleted += sa[2]
    assert deleted > 0


def test_generic_grid_uses_literal_path():
    """Velocity grids that are not the canonical V/G vectors (here: Vx's y-grid
    shifted by a hair) switch the whole advection to the literal bisection code."""
    gr = make_grids(7, 3, uniform=False, stretch=0.2)
    xi_vel = [list(g) for g in gr.xi_vel]
    xi_vel[0][1] = xi_vel[0][1] + 1e-9          # Vx's y-grid no longer equals Vz's y-grid
    xi_vel = tuple(tuple(g) for g in xi_vel)
    o = Oracle(gr.xvi, gr.xci, xi_vel, 16, False); e = Emul(gr.xvi, gr.xci, xi_vel, 16, False)
    assert not e.fast
    co, idx = o.init_particles(8, 9)
    V = stream_velocity(gr)
    dt = cfl_dt(gr, V, 0.8)
    a = [x.copy() for x in co]; b = [x.copy() for x in co]
    o.advect(a, idx, 1, 0.5, V, dt); e.advect(b, idx, 1, 0.5, V, dt)
    assert eq(a, b)


def test_philox_reference_vector():
    """Philox4x32-10 known-answer test (Random123 kat_vectors: ctr=0,key=0 and the
    pi-digits vector) run through the oracle's jpo_rand3 plumbing indirectly:
    the emulation (product source) and the oracle implement Philox independently
    and must agree on every stream."""
    from oracle.oracle import rand3
    gr = make_grids(4, 3, uniform=True)
    o, e = pair(gr, 8)
    co, idx = o.init_particles(8, 0xDEADBEEFCAFE)
    ce, ie = e.init_particles(8, 0xDEADBEEFCAFE)
    assert eq(co, ce)
    r = rand3(0, 0, 0, 0, 0)
    assert ((r >= 0) & (r < 1)).all()
    # Random123 known answer: philox4x32-10, counter = key = 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    expect0 = ((0x6627e8d5 << 32) | 0xe169c58d) >> 11
    assert r[0] == expect0 * 2.0 ** -53
