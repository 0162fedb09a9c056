"""force_injection! (src/Particles/forced_injection.jl:16-79).

The reference's own tests for this entry point state their expected results exactly (counts, which slots keep their
particles, field values, the multiset of coordinates), so they are transcribed here as known-answer tests:
"Forced injection 2D" (test/test_2D.jl:301-384) and "Forced injection 3D" (test/test_3D.jl:279-333).  They pin the
oracle on CPU; on the GPU the library is checked against the oracle bit for bit and against the same expectations."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.problems import make_grids


def _setup_2d():
    """test/test_2D.jl:302-341: 2 x 2 cells, nxcell 0 -> empty container, a 16-particle circle in cell (1, 1), shifted right."""
    gr = make_grids(2, 2, True)
    S, n_circle = 40, 16
    o = Oracle(gr.xvi, gr.xci, gr.xi_vel, S, True)
    co, idx = o.init_particles(0, 1)
    assert not idx.any() and all(np.isnan(c).all() for c in co)       # nxcell = 0: nothing seeded
    ph = np.zeros_like(co[0])
    th = np.linspace(0.0, 2 * np.pi * (1 - 1 / n_circle), n_circle)
    xcirc, ycirc, shift = 0.25 + 0.12 * np.cos(th), 0.5 + 0.12 * np.sin(th), 0.3
    for ip in range(n_circle):
        co[0][ip, 0, 0] = xcirc[ip] + shift; co[1][ip, 0, 0] = ycirc[ip]; idx[ip, 0, 0] = 1; ph[ip, 0, 0] = 1.0
    pnew = [np.full_like(co[0], np.nan), np.zeros_like(co[0])]         # inactive point: coords (0, 0), isnan -> true
    pnew[0][1:] = 0.0                                                  # only the FIRST entry of a cell is tested
    nxp = 8; nyp = -(-S // nxp)
    for c in range(1, S + 1):
        ix = (c - 1) % nxp + 1; j = (c - 1) // nxp + 1
        pnew[0][c - 1, 0, 0] = 0.05 + 0.4 * ix / (nxp + 1); pnew[1][c - 1, 0, 0] = j / (nyp + 1)
    return gr, o, co, idx, ph, pnew, (xcirc + shift, ycirc), S, n_circle


def _check_2d(idx_before, idx, co, ph, expect, S, n_circle):
    active = idx > 0
    injected = (idx_before == 0) & active; existing = (idx_before > 0) & active
    assert int((idx_before > 0).sum()) == n_circle and int(active.sum()) == S
    assert int(injected.sum()) == S - n_circle and int(existing.sum()) == n_circle
    xe, ye = co[0][existing], co[1][existing]
    a = np.lexsort((ye, xe)); b = np.lexsort((expect[1], expect[0]))
    np.testing.assert_allclose(xe[a], expect[0][b]); np.testing.assert_allclose(ye[a], expect[1][b])
    assert (ph[injected] == 3.0).all() and (ph[existing] == 1.0).all()


def test_oracle_forced_injection_2d_reference_test():
    gr, o, co, idx, ph, pnew, expect, S, n_circle = _setup_2d()
    before = idx.copy()
    assert o.force_injection(co, idx, pnew, [ph], [3.0]) == 0
    _check_2d(before, idx, co, ph, expect, S, n_circle)
    # free slot ip took entry ip of p_new (the reference's counter advances with the slot loop)
    inj = (before == 0) & (idx > 0)
    assert np.array_equal(co[0][inj], pnew[0][inj]) and np.array_equal(co[1][inj], pnew[1][inj])
    # no companion fields (test_2D.jl:372-374)
    co2, idx2 = o.init_particles(0, 1)
    o.force_injection(co2, idx2, pnew, [], [])
    assert int(idx2.sum()) == S
    # nothing to inject anywhere (test_2D.jl:376-383)
    _, _, co3, idx3, _, _, _, _, _ = _setup_2d()
    empty = [np.full_like(co3[0], np.nan), np.zeros_like(co3[0])]
    o.force_injection(co3, idx3, empty, [], [])
    assert int(idx3.sum()) == n_circle


def _setup_3d():
    gr = make_grids(2, 3, True)
    S = 4
    o = Oracle(gr.xvi, gr.xci, gr.xi_vel, S, True)
    co, idx = o.init_particles(0, 1)
    pnew = [np.empty_like(co[0]) for _ in range(3)]
    for c in range(1, S + 1):
        for k in range(1, 3):
            for j in range(1, 3):
                for i in range(1, 3):
                    for d, ijk in enumerate((i, j, k)):
                        pnew[d][c - 1, k - 1, j - 1, i - 1] = 0.1 * c + 0.01 * ijk
    return gr, o, co, idx, pnew, S


def test_oracle_forced_injection_3d_reference_test():
    gr, o, co, idx, pnew, S = _setup_3d()
    ph = np.zeros_like(co[0])
    o.force_injection(co, idx, pnew, [ph], [5.0])
    assert idx.all() and (ph == 5.0).all()
    for d in range(3):
        np.testing.assert_allclose(np.sort(co[d].ravel()), np.sort(pnew[d].ravel()))
    co3, idx3 = o.init_particles(0, 1)
    o.force_injection(co3, idx3, [np.full_like(co[0], np.nan)] * 3, [], [])
    assert not idx3.any()


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [2, 3])
def test_gpu_force_injection_reference_tests(ndim):
    import torch
    import justpic.jl_b200 as J
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if ndim == 2:
        gr, o, co, idx, ph, pnew, expect, S, n_circle = _setup_2d()
        p = J.init_particles(J.CUDABackend, 0, S, 5, *gr.grid_vel, seed=1)
        assert p.nxcell == 0 and int(p.index.sum()) == 0
        for d in range(2): p.coords[d].copy_(dev(co[d]))
        p.index.copy_(dev(idx)); gph, = J.init_cell_arrays(p, 1); gph.copy_(dev(ph))
        before = idx.copy()
        J.force_injection(p, [dev(a) for a in pnew], (gph,), (3.0,))
        o.force_injection(co, idx, pnew, [ph], [3.0])
        _check_2d(before, p.index.cpu().numpy(), [c.cpu().numpy() for c in p.coords], gph.cpu().numpy(), expect, S, n_circle)
        val = 3.0
    else:
        gr, o, co, idx, pnew, S = _setup_3d()
        ph = np.zeros_like(co[0])
        p = J.init_particles(J.CUDABackend, 0, S, 0, *gr.grid_vel, seed=1)
        gph, = J.init_cell_arrays(p, 1)
        J.force_injection(p, [dev(a) for a in pnew], (gph,), (5.0,))
        o.force_injection(co, idx, pnew, [ph], [5.0])
        assert bool(p.index.all()) and bool((gph == 5.0).all())
        val = 5.0
    assert np.array_equal(p.index.cpu().numpy(), idx)
    for d in range(ndim):
        assert np.array_equal(p.coords[d].cpu().numpy(), co[d], equal_nan=True)
    assert np.array_equal(gph.cpu().numpy(), ph) and val in ph
    with pytest.raises(ValueError):
        J.force_injection(p, [dev(a) for a in pnew], (gph,), ())


@pytest.mark.gpu
def test_gpu_force_injection_partial_cells_and_wide_slots():
    """Random occupancy, some cells without input, max_xcell > 64: GPU == oracle bit for bit."""
    import torch
    import justpic.jl_b200 as J
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    rng = np.random.default_rng(3)
    for ndim, n, S in [(2, (9, 7), 24), (3, (5, 4, 3), 70)]:
        gr = make_grids(n, ndim, True)
        o = Oracle(gr.xvi, gr.xci, gr.xi_vel, S, True)
        co, idx = o.init_particles(8, 2)
        p = J.init_particles(J.CUDABackend, 8, S, 4, *gr.grid_vel, seed=2)
        assert np.array_equal(p.index.cpu().numpy(), idx)
        pnew = [rng.random(co[0].shape) for _ in range(ndim)]
        pnew[0][0][rng.random(co[0].shape[1:]) < 0.4] = np.nan
        f1 = rng.random(co[0].shape); f2 = rng.random(co[0].shape)
        g1, g2 = dev(f1), dev(f2)
        J.force_injection(p, [dev(a) for a in pnew], (g1, g2), (7.0, -1.5))
        o.force_injection(co, idx, pnew, [f1, f2], [7.0, -1.5])
        assert np.array_equal(p.index.cpu().numpy(), idx)
        for d in range(ndim):
            assert np.array_equal(p.coords[d].cpu().numpy(), co[d], equal_nan=True)
        assert np.array_equal(g1.cpu().numpy(), f1) and np.array_equal(g2.cpu().numpy(), f2)
