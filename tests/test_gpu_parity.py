"""GPU parity: every C-ABI entry point of libjustpic_sm100a.so against the CPU
oracle on the same seeded inputs.  Bar: bit-exact for masks / slot assignment
AND for fp64 values (both sides use only +,-,*,fma,/,sqrt in the same order);
the stated tolerance for floating point is 1e-12 relative (north_star), kept as
a second, looser assertion so a 1-ulp library difference is reported as such."""
import math

import numpy as np
import pytest
import torch

from oracle.oracle import Oracle
from tests.problems import (centre_field_linear, cfl_dt, make_grids, rotation_velocity, stream_velocity,
                            vertex_field_linear)

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def jp():
    import justpic.jl_b200 as J
    return J


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def assert_same(gpu, ref, what):
    g = host(gpu) if isinstance(gpu, torch.Tensor) else gpu
    if g.dtype == np.uint8:
        assert np.array_equal(g, ref), f"{what}: masks differ at {int((g != ref).sum())} slots"
        return
    assert np.array_equal(np.isnan(g), np.isnan(ref)), f"{what}: NaN pattern differs"
    ok = np.isfinite(ref)
    assert np.array_equal(np.isinf(g), np.isinf(ref)), f"{what}: Inf pattern differs"
    np.testing.assert_allclose(g[ok], ref[ok], rtol=RTOL, atol=0, err_msg=f"{what}: beyond the stated 1e-12 tolerance")
    nbad = int((g[ok] != ref[ok]).sum())
    assert nbad == 0, f"{what}: within 1e-12 but not bit-exact at {nbad} entries"


def assert_close(gpu, ref, what, rtol=RTOL):
    """Stated floating-point tolerance of the north_star: 1e-12 relative (to the field's scale
    for entries near zero)."""
    g = host(gpu) if isinstance(gpu, torch.Tensor) else gpu
    assert np.array_equal(np.isnan(g), np.isnan(ref)), f"{what}: NaN pattern differs"
    ok = np.isfinite(ref)
    scale = float(np.abs(ref[ok]).max()) if ok.any() else 1.0
    np.testing.assert_allclose(g[ok], ref[ok], rtol=rtol, atol=rtol * scale, err_msg=what)


class Twin:
    """The same problem on the GPU (public API) and in the oracle."""

    def __init__(self, ndim, n, uniform=True, stretch=0.3, nxcell=12, max_xcell=24, min_xcell=8, seed=7, exact=False):
        J = jp()
        self.gr = gr = make_grids(n, ndim, uniform=uniform, stretch=stretch, exact=exact)
        grids = gr.grid_vel if uniform else gr.xi_vel
        self.p = J.init_particles(J.CUDABackend, nxcell, max_xcell, min_xcell, *grids, seed=seed)
        self.o = Oracle(gr.xvi, gr.xci, gr.xi_vel, self.p.max_xcell, uniform)
        self.co, self.idx = self.o.init_particles(nxcell, seed)
        self.min_xcell = min_xcell
        self.seed = seed

    def check_state(self, what, gargs=(), oargs=()):
        for d in range(self.gr.ndim):
            assert_same(self.p.coords[d], self.co[d], f"{what}: coords[{d}]")
        assert_same(self.p.index, self.idx, f"{what}: index")
        for i, (a, b) in enumerate(zip(gargs, oargs)):
            assert_same(a, b, f"{what}: args[{i}]")


GRIDS = [
    (2, 24, True), (2, (19, 33), False), (3, 10, True), (3, (9, 7, 12), False),
]
ids = lambda g: f"{g[0]}D-{g[1]}-{'range' if g[2] else 'vector'}"


@pytest.mark.parametrize("g", GRIDS, ids=ids)
def test_init_particles(g):
    t = Twin(*g)
    t.check_state("init_particles")
    assert t.p.np == t.p.max_xcell * int(np.prod(t.gr.n))


@pytest.mark.parametrize("g", GRIDS, ids=ids)
@pytest.mark.parametrize("method", ["euler", "rk2", "rk2_23", "rk4"])
def test_advection(g, method):
    J = jp()
    t = Twin(*g)
    V = stream_velocity(t.gr)
    Vd = [dev(v) for v in V]
    m = {"euler": (J.Euler(), 0, 0.0), "rk2": (J.RungeKutta2(), 1, 0.5), "rk2_23": (J.RungeKutta2(2 / 3), 1, 2 / 3),
         "rk4": (J.RungeKutta4(), 2, 0.0)}[method]
    for cfl in (0.5, 1.6):
        dt = cfl_dt(t.gr, V, cfl)
        J.advection(t.p, m[0], Vd, dt)
        t.o.advect(t.co, t.idx, m[1], m[2], V, dt)
        t.check_state(f"advection {method} cfl {cfl}")


@pytest.mark.parametrize("g", [(2, 32, True), (3, 16, True), (3, (32, 8, 16), True)], ids=ids)
@pytest.mark.parametrize("method", ["euler", "rk2", "rk2_23", "rk4"])
@pytest.mark.parametrize("affine", [2, 1, 0])
def test_advection_affine_grids(g, method, affine):
    """Power-of-two grids: vertex vectors are exactly affine (level 1); with Julia range() centres the
    ghosted-centre vectors are too (level 2).  The tiled kernel regenerates such coordinates as
    fma(i, dx, x0) or looks them up (level 0); all must equal the oracle bit for bit."""
    J = jp()
    t = Twin(*g, exact=(affine == 2))
    V = stream_velocity(t.gr)
    Vd = [dev(v) for v in V]
    m = {"euler": (J.Euler(), 0, 0.0), "rk2": (J.RungeKutta2(), 1, 0.5), "rk2_23": (J.RungeKutta2(2 / 3), 1, 2 / 3),
         "rk4": (J.RungeKutta4(), 2, 0.0)}[method]
    for cfl in (0.5, 1.6):
        dt = cfl_dt(t.gr, V, cfl)
        J.advection(t.p, m[0], Vd, dt, affine=affine > 0)
        assert J.advect_affine_level(t.p) == affine
        t.o.advect(t.co, t.idx, m[1], m[2], V, dt)
        t.check_state(f"advection {method} cfl {cfl} affine={affine}")
        J.move_particles(t.p); t.o.move(t.co, t.idx, [])


@pytest.mark.parametrize("g", GRIDS, ids=ids)
def test_interpolations(g):
    J = jp()
    t = Twin(*g)
    gr = t.gr
    T = vertex_field_linear(gr) + 0.25 * np.sin(7 * vertex_field_linear(gr, 0))
    Tc = centre_field_linear(gr) ** 2 + centre_field_linear(gr, 0)
    Td, Tcd = dev(T), dev(Tc)
    pT, = J.init_cell_arrays(t.p, 1)
    opT = np.zeros_like(t.co[0])
    J.grid2particle(pT, Td, t.p); t.o.grid2particle(t.co, t.idx, opT, T)
    assert_same(pT, opT, "grid2particle")
    # reference property: linear field -> pT ≈ coordinate
    lin = vertex_field_linear(gr)
    pL, = J.init_cell_arrays(t.p, 1)
    J.grid2particle(pL, dev(lin), t.p)
    live = t.idx > 0
    np.testing.assert_allclose(host(pL)[live], t.co[-1][live], rtol=math.sqrt(np.finfo(float).eps))
    T2 = torch.empty_like(Td); oT2 = np.empty_like(T)
    J.particle2grid(T2, pT, t.p, mode="exact"); t.o.particle2grid(t.co, t.idx, oT2, opT)
    assert_same(T2, oT2, "particle2grid (exact mode)")
    T3 = torch.empty_like(Td)
    J.particle2grid(T3, pT, t.p, mode="twopass")
    assert_close(T3, oT2, "particle2grid (two-pass mode)")
    T5 = torch.empty_like(Td)
    J.particle2grid(T5, pT, t.p, mode="twopass_fastw")
    assert_close(T5, oT2, "particle2grid (two-pass, fast weights)")
    T4 = torch.empty_like(Td)
    J.particle2grid(T4, pT, t.p, mode="twopass")
    assert torch.equal(T3, T4), "two-pass particle2grid must be deterministic run to run"
    J.centroid2particle(pT, Tcd, t.p); t.o.centroid2particle(t.co, opT, Tc)
    assert_same(pT, opT, "centroid2particle")
    Tc2 = torch.empty_like(Tcd); oTc2 = np.empty_like(Tc)
    J.particle2centroid(Tc2, pT, t.p); t.o.particle2centroid(t.co, oTc2, opT)
    assert_same(Tc2, oTc2, "particle2centroid")


@pytest.mark.parametrize("g", GRIDS, ids=ids)
@pytest.mark.parametrize("K", [2, 5])
def test_phase_ratios_center(g, K):
    J = jp()
    t = Twin(*g)
    rng = np.random.default_rng(K)
    ph = rng.integers(1, K + 1, size=t.idx.shape).astype(np.float64)
    pr = J.PhaseRatios(J.CUDABackend, K, t.gr.n)
    J.phase_ratios_center(pr, t.p, dev(ph))
    ratios = np.zeros(t.o.cell_shape(K))
    t.o.phase_ratios_center(t.co, ratios, ph, K)
    assert_same(pr.center, ratios, "phase_ratios_center")
    np.testing.assert_allclose(host(pr.center).sum(axis=0), 1.0, rtol=1e-14)


@pytest.mark.parametrize("move_mode", ["auto", "direct"])
@pytest.mark.parametrize("g", [(2, 24, True), (2, (19, 33), False), (3, 10, True), (3, (9, 7, 12), False)], ids=ids)
def test_move_policy_compact(g, move_mode):
    """JP_MOVE_POLICY_COMPACT (opt-in, not reference behaviour): bit-exact against the oracle run with the same
    rule (free-slot search restarts at slot 0 for every migrant), and equivalent to the reference policy
    in what a cell CONTAINS: same particles in the same cells (sorted per-cell coordinate multisets) whenever
    nothing was dropped."""
    J = jp()
    t = Twin(*g, nxcell=12, max_xcell=24, min_xcell=8)
    ref = Twin(*g, nxcell=12, max_xcell=24, min_xcell=8)            # same seed: identical start, reference policy
    V = stream_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(t.gr, V, 0.9)
    pT, = J.init_cell_arrays(t.p, 1); J.grid2particle(pT, dev(vertex_field_linear(t.gr)), t.p)
    opT = host(pT).copy(); rpT = host(pT).copy()
    try:
        for it in range(6):
            J.advection(t.p, J.RungeKutta2(), Vd, dt)
            t.o.advect(t.co, t.idx, 1, 0.5, V, dt); ref.o.advect(ref.co, ref.idx, 1, 0.5, V, dt)
            J.move_particles(t.p, (pT,), mode=move_mode, policy="compact")
            Oracle.set_move_policy(True); st = t.o.move(t.co, t.idx, [opT]); Oracle.set_move_policy(False)
            rst = ref.o.move(ref.co, ref.idx, [rpT])
            t.check_state(f"step {it} move_particles[{move_mode}, compact]", (pT,), (opT,))
            assert J.move_stats(t.p) == st
            if st[1] == 0 and rst[1] == 0:                            # nothing dropped: same content per cell
                for d in range(t.gr.ndim):
                    a = np.sort(np.nan_to_num(t.co[d], nan=np.inf), axis=0); b = np.sort(np.nan_to_num(ref.co[d], nan=np.inf), axis=0)
                    assert np.array_equal(a, b), f"step {it}: cell contents differ between the policies (coords[{d}])"
    finally:
        Oracle.set_move_policy(False)


@pytest.mark.parametrize("handoffs", [False, True], ids=["plain", "handoffs"])
@pytest.mark.parametrize("g", [(2, 24, True), (2, (19, 33), False), (3, 10, True), (3, (9, 7, 12), False), (3, (34, 9, 8), True)], ids=ids)
def test_move_policy_dense(g, handoffs):
    """JP_MOVE_POLICY_DENSE (opt-in, not reference behaviour; "vacate everything, then place"): bit-exact against the oracle's twin
    (jpo_move_dense), with and without the two hand-offs; same particles in the same cells as the reference policy whenever nothing
    was dropped; and the cells end up packed at least as low as under the reference rule."""
    J = jp()
    t = Twin(*g, nxcell=12, max_xcell=24, min_xcell=8)
    ref = Twin(*g, nxcell=12, max_xcell=24, min_xcell=8)
    V = stream_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(t.gr, V, 0.9)
    pT, = J.init_cell_arrays(t.p, 1); J.grid2particle(pT, dev(vertex_field_linear(t.gr)), t.p)
    opT = host(pT).copy(); rpT = host(pT).copy()
    T = vertex_field_linear(t.gr)
    F = dev(np.zeros_like(T))
    if handoffs:
        J.move_interp_handoff(t.p, Fp=pT)
    try:
        for it in range(6):
            J.advection(t.p, J.RungeKutta2(), Vd, dt, classify=handoffs)
            t.o.advect(t.co, t.idx, 1, 0.5, V, dt); ref.o.advect(ref.co, ref.idx, 1, 0.5, V, dt)
            J.move_particles(t.p, (pT,), policy="dense")
            assert J.last_move_path(t.p) == "plan"
            Oracle.set_move_policy("dense"); st = t.o.move(t.co, t.idx, [opT]); Oracle.set_move_policy(False)
            rst = ref.o.move(ref.co, ref.idx, [rpT])
            t.check_state(f"step {it} move_particles[dense]", (pT,), (opT,))
            assert J.move_stats(t.p) == st
            J.particle2grid(F, pT, t.p)                                   # (uses the hand-off's cell sums when enabled)
            assert J.last_interp_handoff(t.p)[0] == handoffs
            oF = np.empty_like(T); t.o.particle2grid(t.co, t.idx, oF, opT)
            assert_close(F, oF, f"step {it} particle2grid after a dense move")
            if st[1] == 0 and rst[1] == 0:
                for d in range(t.gr.ndim):
                    a = np.sort(np.nan_to_num(t.co[d], nan=np.inf), axis=0); b = np.sort(np.nan_to_num(ref.co[d], nan=np.inf), axis=0)
                    assert np.array_equal(a, b), f"step {it}: cell contents differ between the policies (coords[{d}])"
                top = lambda idx: (idx * (np.arange(idx.shape[0]) + 1).reshape((-1,) + (1,) * (idx.ndim - 1))).max(axis=0)
                assert top(t.idx).sum() <= top(ref.idx).sum(), "dense policy: cells are packed at least as low as under the reference rule"
        with pytest.raises(Exception):
            J.move_particles(t.p, (pT,), mode="direct", policy="dense")    # planned path only
    finally:
        Oracle.set_move_policy(False)
        J.move_particles(t.p, (pT,), policy="reference")


@pytest.mark.parametrize("move_mode", ["auto", "direct"])
@pytest.mark.parametrize("g", GRIDS + [(2, 17, True), (3, (7, 5, 6), True), (2, (40, 9), True)], ids=ids)
def test_trajectory_advect_move_inject(g, move_mode):
    """L2 protocol: coupled steps; any divergence shows up at the first differing call.
    Both move implementations (planned sweeps + streaming payload passes, and the direct
    literal sweeps) must reproduce the reference's slot assignment bit for bit."""
    J = jp()
    tight = g in [(2, 17, True)]
    t = Twin(*g, nxcell=12, max_xcell=12 if tight else 24, min_xcell=6 if tight else 8)
    gr = t.gr
    V = stream_velocity(gr)
    Vd = [dev(v) for v in V]
    dt = cfl_dt(gr, V, 0.9)
    T = vertex_field_linear(gr)
    pT, ph = J.init_cell_arrays(t.p, 2)
    J.grid2particle(pT, dev(T), t.p)
    opT = np.zeros_like(t.co[0]); t.o.grid2particle(t.co, t.idx, opT, T)
    oph = np.where(t.idx > 0, 1.0 + (t.co[0] < t.co[-1]), 0.0)
    ph.copy_(dev(oph))
    methods = [(J.RungeKutta2(), 1, 0.5), (J.RungeKutta4(), 2, 0.0), (J.RungeKutta2(2 / 3), 1, 2 / 3), (J.Euler(), 0, 0.0)]
    for it in range(8):
        m = methods[it % 4]
        J.advection(t.p, m[0], Vd, dt); t.o.advect(t.co, t.idx, m[1], m[2], V, dt)
        t.check_state(f"step {it} advection")
        J.move_particles(t.p, (pT, ph), mode=move_mode); st = t.o.move(t.co, t.idx, [opT, oph])
        t.check_state(f"step {it} move_particles[{move_mode}]", (pT, ph), (opT, oph))
        assert J.move_stats(t.p) == st
        assert J.last_move_path(t.p) == ("plan" if move_mode == "auto" else "direct")
        J.inject_particles(t.p, (pT, ph), step=it); inj = t.o.inject(t.co, t.idx, [opT, oph], t.min_xcell, t.seed, it)
        t.check_state(f"step {it} inject_particles", (pT, ph), (opT, oph))
        assert J.inject_stats(t.p) == inj
    Tg = torch.empty_like(dev(T)); oT = np.empty_like(T)
    J.particle2grid(Tg, pT, t.p, mode="exact"); t.o.particle2grid(t.co, t.idx, oT, opT)
    assert_same(Tg, oT, "final particle2grid (exact)")
    J.particle2grid(Tg, pT, t.p, mode="twopass")
    assert_close(Tg, oT, "final particle2grid (two-pass)")


@pytest.mark.parametrize("ndim", [2, 3])
def test_move_many_leavers_full_occupancy_word(ndim):
    """64 slots, ~50 particles per cell, large CFL: cells with more than 24 / 36 leavers
    (several packed code / result words per cell) and destinations that fill up."""
    J = jp()
    t = Twin(ndim, 12 if ndim == 2 else (6, 5, 7), True, nxcell=48, max_xcell=64, min_xcell=16, seed=11)
    V = stream_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(t.gr, V, 0.98)
    pT, = J.init_cell_arrays(t.p, 1)
    opT = np.where(t.idx > 0, t.co[0] * 3.0, 0.0); pT.copy_(dev(opT))
    for it in range(5):
        J.advection(t.p, J.RungeKutta2(), Vd, dt); t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
        J.move_particles(t.p, (pT,)); st = t.o.move(t.co, t.idx, [opT])
        t.check_state(f"many leavers step {it}", (pT,), (opT,))
        assert J.move_stats(t.p) == st and J.last_move_path(t.p) == "plan"


@pytest.mark.parametrize("ndim", [2, 3])
def test_ties_nan_inf_and_clean(ndim):
    J = jp()
    t = Twin(ndim, 6, True, nxcell=8, max_xcell=12, min_xcell=8, seed=5)
    gr = t.gr
    rng = np.random.default_rng(1)
    live = np.argwhere(t.idx > 0)
    for tt in live[rng.choice(len(live), size=len(live) // 3, replace=False)]:
        for d in range(ndim):
            if rng.random() < 0.6:
                cell = tt[ndim - d]
                cand = [gr.xvi[d][cell], gr.xvi[d][cell + 1], gr.xci[d][cell], np.nan, np.inf, -1e-3, 1.001,
                        np.nextafter(gr.xvi[d][cell + 1], 2.0)]
                t.co[d][tuple(tt)] = cand[rng.integers(len(cand))]
    for d in range(ndim):
        t.p.coords[d].copy_(dev(t.co[d]))
    pT, = J.init_cell_arrays(t.p, 1)
    opT = np.where(t.idx > 0, rng.random(t.idx.shape), 0.0)
    pT.copy_(dev(opT))
    # clean on a copy
    c2 = [a.copy() for a in t.co]; i2 = t.idx.copy(); a2 = opT.copy()
    t.o.clean(c2, i2, [a2])
    import copy
    pc = [c.clone() for c in t.p.coords]; pi = t.p.index.clone(); pa = pT.clone()
    J.clean_particles(t.p, None, (pT,))
    for d in range(ndim):
        assert_same(t.p.coords[d], c2[d], "clean coords")
    assert_same(t.p.index, i2, "clean index"); assert_same(pT, a2, "clean args")
    for d in range(ndim):
        t.p.coords[d].copy_(pc[d])
    t.p.index.copy_(pi); pT.copy_(pa)
    # move straight from the tie positions (both implementations), on copies
    for mode in ("auto", "direct"):
        c3 = [a.copy() for a in t.co]; i3 = t.idx.copy(); a3 = opT.copy()
        st = t.o.move(c3, i3, [a3])
        J.move_particles(t.p, (pT,), mode=mode)
        for d in range(ndim):
            assert_same(t.p.coords[d], c3[d], f"tie move[{mode}] coords")
        assert_same(t.p.index, i3, f"tie move[{mode}] index"); assert_same(pT, a3, f"tie move[{mode}] args")
        assert J.move_stats(t.p) == st and st[2] > 0
        assert J.last_move_path(t.p) == "direct"          # particles exactly on faces: literal sweeps
        for d in range(ndim):
            t.p.coords[d].copy_(pc[d])
        t.p.index.copy_(pi); pT.copy_(pa)
    V = stream_velocity(gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(gr, V, 0.5)
    for it, m in enumerate([(J.RungeKutta2(), 1, 0.5), (J.RungeKutta4(), 2, 0.0), (J.Euler(), 0, 0.0)]):
        J.advection(t.p, m[0], Vd, dt); t.o.advect(t.co, t.idx, m[1], m[2], V, dt)
        t.check_state(f"ties advection {it}")
        J.move_particles(t.p, (pT,)); st = t.o.move(t.co, t.idx, [opT])
        t.check_state(f"ties move {it}", (pT,), (opT,))
        assert J.move_stats(t.p) == st
        J.inject_particles(t.p, (pT,), step=it); t.o.inject(t.co, t.idx, [opT], 8, 5, it)
        t.check_state(f"ties inject {it}", (pT,), (opT,))


def test_rotating_circle_rk4_inject_cell_assignment():
    """BASELINE config 2 in miniature (scripts/rotating_circle.jl): solid rotation, RK4,
    inject_particles! reseeding; cell assignment (masks) must be bit-exact."""
    J = jp()
    t = Twin(2, 48, True, nxcell=24, max_xcell=48, min_xcell=12, seed=42)
    V = rotation_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = 200.0 * 4          # <= 1 cell per step: the regime in which the reference's colour sweeps are race-free
    ph, = J.init_cell_arrays(t.p, 1)
    r2 = (t.co[0] - 0.5) ** 2 + (t.co[1] - 0.75) ** 2
    oph = np.where(t.idx > 0, 1.0 + (r2 < 0.15 ** 2), 0.0)
    ph.copy_(dev(oph))
    for it in range(6):
        J.advection(t.p, J.RungeKutta4(), Vd, dt); t.o.advect(t.co, t.idx, 2, 0.0, V, dt)
        J.move_particles(t.p, (ph,)); t.o.move(t.co, t.idx, [oph])
        J.inject_particles(t.p, (ph,), step=it); t.o.inject(t.co, t.idx, [oph], 12, 42, it)
        t.check_state(f"rotating circle step {it}", (ph,), (oph,))


def test_api_errors():
    J = jp()
    t = Twin(2, 8, True)
    V = [dev(v) for v in stream_velocity(t.gr)]
    with pytest.raises(ValueError):
        J.RungeKutta2(1.1)
    with pytest.raises(ValueError):
        J.advection(t.p, J.Euler(), V[:1], 0.1)
    with pytest.raises(ValueError):
        J.grid2particle(torch.zeros(3, device="cuda", dtype=torch.float64), V[0], t.p)
    with pytest.raises(ValueError):
        J.init_particles(J.CUDABackend, 12, 24, 8)


# ------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("cfg", ["cfg1_2d_256", "cfg2_2d_512_rk4", "cfg3_3d_128", "cfg4_3d_256"])
def test_full_size_invariants(cfg):
    """BASELINE.json sizes: the oracle is too slow, so size-independent properties:
    count conservation (live = initial - dropped - deleted + injected), every live
    particle inside its cell, dead slots NaN, linear field reproduced by g2p, p2g of
    a constant field is that constant, phase ratios sum to 1, idempotent move."""
    J = jp()
    ndim, n = {"cfg1_2d_256": (2, 256), "cfg2_2d_512_rk4": (2, 512), "cfg3_3d_128": (3, 128), "cfg4_3d_256": (3, 256)}[cfg]
    if cfg == "cfg4_3d_256" and torch.cuda.mem_get_info()[1] < 100e9:
        pytest.skip("needs ~80 GB of device memory")
    gr = make_grids(n, ndim, True)
    p = J.init_particles(J.CUDABackend, 24, 48, 12, *gr.grid_vel, seed=42)
    rk4 = cfg == "cfg2_2d_512_rk4"                       # BASELINE configs[1]: rotation field, RK4, inject reseeding
    V = rotation_velocity(gr) if rk4 else stream_velocity(gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(gr, V, 0.75 if ndim == 2 else 0.5)
    method = J.RungeKutta4() if rk4 else J.RungeKutta2()
    T = dev(vertex_field_linear(gr))
    pT, pc = J.init_cell_arrays(p, 2)
    J.grid2particle(pT, T, p)
    pc.fill_(3.25)
    n_live = int(p.index.sum().item())
    assert n_live == 24 * int(np.prod(gr.n))
    for it in range(3):
        J.advection(p, method, Vd, dt)
        J.move_particles(p, (pT, pc))
        moved, dropped, deleted = J.move_stats(p)
        J.inject_particles(p, (pT, pc), step=it)
        inj = J.inject_stats(p)
        n_new = int(p.index.sum().item())
        assert n_new == n_live - dropped - deleted + inj
        assert moved > 0
        n_live = n_new
    live = p.index > 0
    for d in range(ndim):
        shape = [1] * (ndim + 1); shape[ndim - d] = -1
        lo = dev(gr.xvi[d][:-1]).reshape(shape); hi = dev(gr.xvi[d][1:]).reshape(shape)
        c = p.coords[d]
        assert bool((((c >= lo) & (c <= hi)) | ~live).all())
        assert bool((torch.isnan(c) == ~live).all())
    # idempotence: a second move changes nothing (checksums of position-weighted sums: no 20 GB clones at 256^3)
    def digest():
        w = torch.arange(p.index.shape[0], device="cuda", dtype=torch.float64).reshape(-1, *([1] * ndim)) + 1.0
        return [float((torch.nan_to_num(c, nan=-7.0) * w).sum()) for c in p.coords] + [float((p.index.double() * w).sum())]
    before = digest()
    J.move_particles(p, (pT, pc))
    assert J.move_stats(p) == (0, 0, 0)
    assert digest() == before
    # constant field survives p2g exactly up to rounding of the weighted mean
    F = torch.empty_like(T)
    J.particle2grid(F, pc, p)
    assert bool(torch.allclose(F, torch.full_like(F, 3.25), rtol=1e-13, atol=0))
    pL, = J.init_cell_arrays(p, 1)
    J.grid2particle(pL, T, p)
    assert bool(torch.allclose(pL[live], p.coords[-1][live], rtol=1e-8, atol=1e-12))
    K = 3
    ph = torch.where(live, 1.0 + torch.floor(p.coords[0].nan_to_num() * K).clamp(0, K - 1), torch.zeros_like(pL))
    pr = J.PhaseRatios(J.CUDABackend, K, gr.n)
    J.phase_ratios_center(pr, p, ph)
    s = pr.center.sum(dim=0)
    assert bool(torch.allclose(s, torch.ones_like(s), rtol=1e-13))


def test_unsupported_and_invalid_arguments():
    """Error behaviour across the boundary: status codes -> exceptions, nothing crashes."""
    J = jp()
    gr = make_grids(6, 2, True)
    with pytest.raises(J._cabi.JustPICError):            # max_xcell > JP_MAX_SLOTS_WIDE (1024): JP_ERR_UNSUPPORTED
        J.init_particles(J.CUDABackend, 8, 1025, 4, *gr.grid_vel)
    t = Twin(2, 6, True)
    many = J.init_cell_arrays(t.p, 17)
    with pytest.raises(ValueError):                      # more than JP_MAX_ARGS fields in one call
        J.move_particles(t.p, many)
    with pytest.raises(ValueError):
        J.particle2grid(torch.zeros(7 * 7, device="cuda", dtype=torch.float64), many[0], t.p, mode="bogus")
    with pytest.raises(ValueError):                      # wrong dtype
        J.grid2particle(many[0].float(), torch.zeros(7 * 7, device="cuda", dtype=torch.float64), t.p)
    pr = J.PhaseRatios(J.CUDABackend, 40, gr.n)          # more than JP_MAX_PHASES
    with pytest.raises(J._cabi.JustPICError):
        J.phase_ratios_center(pr, t.p, many[0])


WIDE = [
    # ndim, n, uniform, nxcell, max_xcell, min_xcell  -- the reference's own sizes above 64 slots
    (2, (11, 9), True, 60, 80, 50),        # test/test_2D.jl:437
    (3, (5, 4, 6), False, 125, 150, 100),  # test/test_3D.jl:369,424 (refined grid variant)
    (3, (6, 5, 4), True, 40, 70, 32),      # second chunk only 6 slots wide
]


@pytest.mark.parametrize("w", WIDE, ids=lambda w: f"{w[0]}D-{w[1]}-{'range' if w[2] else 'vector'}-S{w[4]}")
def test_wide_cells_max_xcell_above_64(w):
    """max_xcell > 64 (JP_MAX_SLOTS): the per-slot kernels run in 64-slot chunks, move_particles! /
    inject_particles!(_phase!) take the literal per-cell kernels, particle2grid! the exact kernel.  Every
    entry point of a coupled run must still match the oracle bit for bit."""
    J = jp()
    ndim, n, uniform, nxcell, max_xcell, min_xcell = w
    t = Twin(ndim, n, uniform=uniform, nxcell=nxcell, max_xcell=max_xcell, min_xcell=min_xcell)
    assert t.p.max_xcell == max_xcell and t.idx.shape[0] == max_xcell
    t.check_state("init_particles")
    gr = t.gr
    V = stream_velocity(gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(gr, V, 0.8)
    T = vertex_field_linear(gr) + 0.25 * np.sin(7 * vertex_field_linear(gr, 0)); Tc = centre_field_linear(gr) ** 2
    Td, Tcd = dev(T), dev(Tc)
    pT, ph, pC = J.init_cell_arrays(t.p, 3)
    opT = np.zeros_like(t.co[0]); opC = np.zeros_like(t.co[0])
    J.grid2particle(pT, Td, t.p); t.o.grid2particle(t.co, t.idx, opT, T)
    assert_same(pT, opT, "grid2particle")
    J.centroid2particle(pC, Tcd, t.p); t.o.centroid2particle(t.co, opC, Tc)
    assert_same(pC, opC, "centroid2particle")
    oph = np.where(t.idx > 0, 1.0 + (t.co[0] < t.co[-1]), 0.0)
    ph.copy_(dev(oph))
    K = 2
    pr = J.PhaseRatios(J.CUDABackend, K, gr.n)
    methods = [(J.RungeKutta2(), 1, 0.5), (J.RungeKutta4(), 2, 0.0), (J.RungeKutta2(2 / 3), 1, 2 / 3), (J.Euler(), 0, 0.0)]
    moved = injected = 0
    for it in range(6):
        m = methods[it % 4]
        if it == 4:
            J.advection_LinP(t.p, m[0], Vd, dt); assert t.o.advect_interp(t.co, t.idx, m[1], m[2], V, dt, 1) == 0
        elif it == 5:
            J.advection_MQS(t.p, m[0], Vd, dt); assert t.o.advect_interp(t.co, t.idx, m[1], m[2], V, dt, 2) == 0
        else:
            J.advection(t.p, m[0], Vd, dt); t.o.advect(t.co, t.idx, m[1], m[2], V, dt)
        t.check_state(f"step {it} advection")
        J.move_particles(t.p, (pT, ph, pC)); st = t.o.move(t.co, t.idx, [opT, oph, opC])
        t.check_state(f"step {it} move_particles", (pT, ph, pC), (opT, oph, opC))
        assert J.move_stats(t.p) == st and J.last_move_path(t.p) == "direct"
        moved += st[0]
        if it % 2 == 0:
            J.inject_particles(t.p, (pT, ph, pC), step=it); inj = t.o.inject(t.co, t.idx, [opT, oph, opC], min_xcell, t.seed, it)
        else:
            J.inject_particles_phase(t.p, ph, (pT, pC), (Td, Tcd), step=it)
            inj = t.o.inject_phase(t.co, t.idx, oph, [opT, opC], [T, Tc], [0, 1], min_xcell, t.seed, it)
        t.check_state(f"step {it} inject", (pT, ph, pC), (opT, oph, opC))
        assert J.inject_stats(t.p) == inj
        injected += inj
        J.phase_ratios_center(pr, t.p, ph)
        ratios = np.zeros(t.o.cell_shape(K)); t.o.phase_ratios_center(t.co, ratios, oph, K)
        assert_same(pr.center, ratios, f"step {it} phase_ratios_center")
    assert moved > 0 and injected > 0
    Tg = torch.empty_like(Td); oT = np.empty_like(T)
    t.o.particle2grid(t.co, t.idx, oT, opT)
    for mode in ("exact", "twopass", "twopass_fastw"):
        J.particle2grid(Tg, pT, t.p, mode=mode)
        assert_same(Tg, oT, f"particle2grid[{mode}] (wide cells use the exact kernel)")
    Tc2 = torch.empty_like(Tcd); oTc2 = np.empty_like(Tc)
    J.particle2centroid(Tc2, pT, t.p); t.o.particle2centroid(t.co, oTc2, opT)
    assert_same(Tc2, oTc2, "particle2centroid")
    T0 = T * 0.5
    J.grid2particle_flip(pT, None, Td, dev(T0), t.p, alpha=0.25); t.o.grid2particle_flip(t.co, t.idx, opT, T, T0, 0.25)
    assert_same(pT, opT, "grid2particle_flip")
    # push a few particles out of their cells, then clean_particles!
    sh = t.co[0].copy(); live = t.idx > 0
    sh[live] += np.where(np.arange(int(live.sum())) % 7 == 0, 1.5 * float(gr.xvi[0][1] - gr.xvi[0][0]), 0.0)
    t.co[0][:] = sh; t.p.coords[0].copy_(dev(sh))
    J.clean_particles(t.p, None, (pT, ph, pC)); t.o.clean(t.co, t.idx, [opT, oph, opC])
    t.check_state("clean_particles", (pT, ph, pC), (opT, oph, opC))
    # checkpoint layout round trip
    h = J.Array(t.p); p2 = J.CuArray(h)
    for d in range(ndim):
        assert torch.equal(torch.nan_to_num(p2.coords[d], nan=-1.0), torch.nan_to_num(t.p.coords[d], nan=-1.0))
    assert torch.equal(p2.index, t.p.index)
