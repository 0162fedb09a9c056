"""inject_particles_phase! (src/Particles/injection.jl:146-325).
CPU: properties of the oracle restatement (every quadrant ends up with >= cld(min_xcell, 2^N)
particles when slots allow, new particles sit strictly inside their quadrant, phases are copied
from existing particles, interpolated fields stay within the stencil extrema).
GPU: the CUDA path against the oracle, bit for bit, inside the reference's own time loop
(test/test_2D.jl:468-474: advection! -> move_particles! -> inject_particles_phase! -> update_phase_ratios!)."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from tests.problems import centre_field_linear, cfl_dt, make_grids, stream_velocity, vertex_field_linear


def _thin(o, co, idx, arrs, rng, keep=0.35):
    """Kill a random subset so that many quadrants become deficient."""
    kill = (rng.random(idx.shape) > keep) & (idx > 0)
    idx[kill] = 0
    for a in list(co) + list(arrs):
        a[kill] = np.nan


def _quadrant_counts(gr, co, idx):
    N = gr.ndim
    counts = np.zeros((2 ** N, *idx.shape[1:]), dtype=np.int64)
    shape = idx.shape[1:]
    for q in range(2 ** N):
        inq = idx > 0
        for d in range(N):
            xv = gr.xvi[d]
            dq = (xv[1] - xv[0]) / 2 if gr.uniform else np.diff(xv) / 2
            lo = xv[:-1] + dq * ((q >> d) & 1)
            sh = [1] * (N + 1); sh[N - d] = len(lo)
            lo_b = np.reshape(lo, sh); dq_b = np.reshape(dq, sh) if not np.isscalar(dq) else dq
            with np.errstate(invalid="ignore"):
                inq &= (lo_b < co[d]) & (co[d] < lo_b + dq_b)
        counts[q] = inq.sum(axis=0)
    return counts


@pytest.mark.parametrize("ndim,n,uniform", [(2, (9, 7), True), (2, (6, 8), False), (3, (5, 4, 6), True), (3, (4, 5, 3), False)])
def test_oracle_inject_phase_properties(ndim, n, uniform):
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.3)
    S, min_xcell, K = 24, 8, 3
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    co, idx = o.init_particles(12, 3)
    rng = np.random.default_rng(1)
    ph = np.where(idx > 0, rng.integers(1, K + 1, size=idx.shape), np.nan).astype(np.float64)
    pT = np.full(idx.shape, np.nan); pC = np.full(idx.shape, np.nan)
    _thin(o, co, idx, [ph, pT, pC], rng)
    Tv = vertex_field_linear(gr) + 0.1 * np.sin(5 * vertex_field_linear(gr, 0))
    Tc = centre_field_linear(gr) ** 2
    before = idx.copy()
    ninj = o.inject_phase(co, idx, ph, [pT, pC], [Tv, Tc], [0, 1], min_xcell, 9, 0)
    new = (idx > 0) & (before == 0)
    assert ninj == int(new.sum()) > 0
    assert np.all(idx[before > 0] > 0)                                     # nothing removed
    assert np.all(np.isin(ph[new], np.arange(1, K + 1)))                   # phases copied from live particles
    assert np.all(np.isfinite(pT[new])) and np.all(np.isfinite(pC[new]))
    assert Tv.min() <= pT[new].min() and pT[new].max() <= Tv.max()         # clamped to stencil extrema
    assert Tc.min() <= pC[new].min() and pC[new].max() <= Tc.max()
    counts = _quadrant_counts(gr, co, idx)
    min_xq = math.ceil(min_xcell / 2 ** ndim)
    full = (idx > 0).sum(axis=0) == S
    assert np.all((counts >= min_xq) | full[None])                          # every quadrant topped up unless the cell is full


GPU_CASES = [(2, (24, 17), True), (2, (19, 33), False), (3, (10, 9, 12), True), (3, (9, 7, 12), False)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GPU_CASES, ids=lambda c: f"{c[0]}D-{'x'.join(map(str, c[1]))}-{'range' if c[2] else 'vector'}")
def test_gpu_inject_particles_phase_loop(case):
    import torch
    import justpic.jl_b200 as J
    ndim, n, uniform = case
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.3)
    grids = gr.grid_vel if uniform else gr.xi_vel
    S, min_xcell, K, seed = 24, 10, 3, 11
    p = J.init_particles(J.CUDABackend, 12, S, min_xcell, *grids, seed=seed)
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    co, idx = o.init_particles(12, seed)
    rng = np.random.default_rng(2)
    oph = np.where(idx > 0, rng.integers(1, K + 1, size=idx.shape), np.nan).astype(np.float64)
    opT = np.full(idx.shape, np.nan); opC = np.full(idx.shape, np.nan)
    _thin(o, co, idx, [oph, opT, opC], rng, keep=0.5)
    Tv = np.ascontiguousarray(vertex_field_linear(gr) + 0.1 * np.sin(5 * vertex_field_linear(gr, 0)))
    Tc = np.ascontiguousarray(centre_field_linear(gr) ** 2)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for d in range(ndim):
        p.coords[d].copy_(dev(co[d]))
    p.index.copy_(dev(idx))
    ph, pT, pC = dev(oph), dev(opT), dev(opC)
    Tvd, Tcd = dev(Tv), dev(Tc)
    V = stream_velocity(gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(gr, V, 0.7)
    pr = J.PhaseRatios(J.CUDABackend, K, gr.n)

    def same(a, b, what):
        a = a.cpu().numpy()
        assert np.array_equal(a, b, equal_nan=True), f"{what}: {int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())} entries differ"

    total = 0
    for it in range(3):
        J.advection(p, J.RungeKutta2(), Vd, dt); o.advect(co, idx, 1, 0.5, V, dt)
        J.move_particles(p, (ph, pT, pC)); o.move(co, idx, [oph, opT, opC])
        J.inject_particles_phase(p, ph, (pT, pC), (Tvd, Tcd), step=it)
        ninj = o.inject_phase(co, idx, oph, [opT, opC], [Tv, Tc], [0, 1], min_xcell, seed, it)
        assert J.inject_stats(p) == ninj
        total += ninj
        same(p.index, idx, f"step {it}: index")
        for d in range(ndim):
            same(p.coords[d], co[d], f"step {it}: coords[{d}]")
        same(ph, oph, f"step {it}: phases"); same(pT, opT, f"step {it}: vertex-interpolated field"); same(pC, opC, f"step {it}: centre-interpolated field")
        J.update_phase_ratios(pr, p, ph)
        s = pr.center.cpu().numpy().sum(axis=0)
        np.testing.assert_allclose(s[np.isfinite(s)], 1.0, rtol=1e-13)     # test_2D.jl:476-479
    assert total > 0


@pytest.mark.gpu
def test_gpu_inject_particles_phase_no_fields():
    """The reference's own call: inject_particles_phase!(particles, phases, (), ()) (test/test_2D.jl:472)."""
    import torch
    import justpic.jl_b200 as J
    gr = make_grids(12, 2, True)
    p = J.init_particles(J.CUDABackend, 8, 16, 8, *gr.grid_vel, seed=3)
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, 16, True)
    co, idx = o.init_particles(8, 3)
    oph = np.where(idx > 0, 1.0 + (co[0] < co[1]), np.nan)
    rng = np.random.default_rng(0)
    _thin(o, co, idx, [oph], rng, keep=0.4)
    for d in range(2):
        p.coords[d].copy_(torch.from_numpy(co[d]))
    p.index.copy_(torch.from_numpy(idx))
    ph = torch.from_numpy(oph).cuda()
    J.inject_particles_phase(p, ph, (), (), step=0)
    o.inject_phase(co, idx, oph, [], [], [], 8, 3, 0)
    assert np.array_equal(p.index.cpu().numpy(), idx)
    assert np.array_equal(ph.cpu().numpy(), oph, equal_nan=True)
    with pytest.raises(ValueError):
        J.inject_particles_phase(p, ph, (ph,), (), step=1)
