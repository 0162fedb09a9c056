"""Pins the CPU oracle against the reference's own known-answer and property
tests.  CPU only."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from tests.problems import (centre_field_linear, cfl_dt, make_grids, stream_velocity, vertex_field_linear)

KATS = json.loads((Path(__file__).parent / "golden" / "reference_kats.json").read_text())
RTOL = math.sqrt(np.finfo(np.float64).eps)   # Julia's default for `≈`


@pytest.mark.parametrize("k", KATS["first_stage"], ids=lambda k: k["src"])
def test_first_stage_kat(k):
    out = O.first_stage(0 if k["scheme"] == "euler" else 1, k["alpha"], k["dt"], k["v"], k["p"])
    if k["exact"]:
        assert out.tolist() == k["expect"]
    else:
        np.testing.assert_allclose(out, k["expect"], rtol=RTOL)


@pytest.mark.parametrize("k", KATS["second_stage"], ids=lambda k: k["src"])
def test_second_stage_kat(k):
    out = O.second_stage(k["alpha"], k["dt"], k["v0"], k["v1"], k["p"])
    np.testing.assert_allclose(out, k["expect"], rtol=RTOL)


@pytest.mark.parametrize("k", KATS["lerp"], ids=lambda k: k["src"])
def test_lerp_kat(k):
    assert O.lerp(k["v"], k["t"]) == k["expect"]


def test_rk2_alpha_validation():
    # test/test_integrators.jl:8-9: RungeKutta2(1.1) / RungeKutta2(-0.1) throw ArgumentError
    from justpic.jl_b200.api import Euler, RungeKutta2
    assert RungeKutta2().alpha == 0.5 and RungeKutta2(2 / 3).alpha == 2 / 3
    with pytest.raises(ValueError):
        RungeKutta2(1.1)
    with pytest.raises(ValueError):
        RungeKutta2(-0.1)
    assert isinstance(Euler(1), Euler) and isinstance(Euler("potato"), Euler)


def _setup(ndim, n, uniform=True, S=12, nxcell=5, stretch=0.3, seed=3):
    gr = make_grids(n, ndim, uniform=uniform, stretch=stretch)
    NQ = 2 ** ndim
    S = max(S, math.ceil(nxcell / NQ) * NQ)      # max_xcell = max(nxcell, max_xcell), particles_utils.jl:151-155
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    coords, index = o.init_particles(nxcell, seed)
    return gr, o, coords, index


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("uniform", [True, False])
def test_init_particles_in_cells_and_quadrants(ndim, uniform):
    # fill_coords_index! (particles_utils.jl:168-194): np_quadrant per quadrant, first nxcell slots live
    gr, o, coords, index = _setup(ndim, 6, uniform, S=20, nxcell=10)
    NQ = 2 ** ndim
    npq = math.ceil(10 / NQ)
    assert index[: npq * NQ].all() and not index[npq * NQ:].any()
    assert all(np.isnan(c[npq * NQ:]).all() for c in coords)
    for d in range(ndim):
        lo = gr.xvi[d][:-1]; hi = gr.xvi[d][1:]
        shape = [1] * (ndim + 1); shape[ndim - d] = -1
        c = coords[d][: npq * NQ]
        assert (c >= lo.reshape(shape)).all() and (c <= hi.reshape(shape)).all()
        # quadrant membership: slot l belongs to quadrant l // npq, bit d = upper half
        mid = 0.5 * (lo + hi)
        for l in range(npq * NQ):
            upper = ((l // npq) >> d) & 1
            if upper:
                assert (c[l] >= mid.reshape(shape[1:]) - 1e-15).all()
            else:
                assert (c[l] <= mid.reshape(shape[1:]) + 1e-15).all()


@pytest.mark.parametrize("ndim", [2, 3])
def test_linear_field_interpolation_properties(ndim):
    # test/test_interpolation_kernels.jl:63-123 (2D), :125-188 (3D)
    n = 4
    gr, o, coords, index = _setup(ndim, n, True, S=5 if ndim == 2 else 8, nxcell=5)
    T = vertex_field_linear(gr)
    Tc = centre_field_linear(gr)
    pT = np.zeros_like(coords[0])
    o.grid2particle(coords, index, pT, T)
    live = index > 0
    np.testing.assert_allclose(pT[live], coords[-1][live], rtol=RTOL)          # pT ≈ coords[N]
    T2 = np.empty_like(T)
    o.particle2grid(coords, index, T2, pT)
    assert np.linalg.norm(T2 - T) / T.size < 1e-1
    o.centroid2particle(coords, pT, Tc)
    interior = (slice(None),) + (1,) * ndim                                     # cell (2,2[,2]) 1-based
    np.testing.assert_allclose(pT[interior][live[interior]], coords[-1][interior][live[interior]], rtol=RTOL)
    Tc2 = np.empty_like(Tc)
    o.particle2centroid(coords, Tc2, pT)
    assert np.linalg.norm(Tc2 - Tc) / Tc.size < 1e-1


@pytest.mark.parametrize("ndim", [2, 3])
def test_phase_ratios_sum_to_one(ndim):
    # test/test_CellArrays.jl:81-113 / :144-181: 5 random phases, centre ratios sum to 1
    gr, o, coords, index = _setup(ndim, 6, True, S=16, nxcell=12)
    rng = np.random.default_rng(0)
    K = 5
    phases = rng.integers(1, K + 1, size=coords[0].shape).astype(np.float64)
    ratios = np.zeros(o.cell_shape(K))
    o.phase_ratios_center(coords, ratios, phases, K)
    np.testing.assert_allclose(ratios.sum(axis=0), 1.0, rtol=1e-14)
    assert (ratios >= 0).all()


@pytest.mark.parametrize("ndim", [2, 3])
def test_move_invariants(ndim):
    # .agents/validation.md-style invariants: count conservation, every live particle strictly in its cell
    gr, o, coords, index = _setup(ndim, 10 if ndim == 2 else 7, True, S=24, nxcell=12)
    V = stream_velocity(gr)
    dt = cfl_dt(gr, V, 0.9)
    pT = np.where(index > 0, coords[0], 0.0)
    n0 = int(index.sum())
    o.advect(coords, index, 1, 0.5, V, dt)
    moved, dropped, deleted = o.move(coords, index, [pT])
    assert int(index.sum()) == n0 - dropped - deleted
    live = index > 0
    for d in range(ndim):
        shape = [1] * (ndim + 1); shape[ndim - d] = -1
        lo = gr.xvi[d][:-1].reshape(shape); hi = gr.xvi[d][1:].reshape(shape)
        c = coords[d]
        assert ((c >= lo) & (c <= hi))[live].all()
        assert np.isnan(c[~live]).all()
    assert not np.isnan(pT[live]).any()      # fields travel with their particle


def test_miniapp_temperature_conservation_2d():
    # test/test_2D.jl:482-538: stream function, RK2(2/3), 25 iterations of
    # p2g -> advect -> move -> inject -> g2p;  |sum(T) - sum(T0)| / sum(T0) < 1e-2
    n = 64
    gr = make_grids(n, 2, True)
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, 48, True)
    coords, index = o.init_particles(24, 42)
    T = vertex_field_linear(gr)
    T0 = T.copy()
    V = stream_velocity(gr)
    dt = cfl_dt(gr, V, 1.0) / 10          # test_2D.jl: dt = min(dx/max|V|) ... scaled
    pT = np.zeros_like(coords[0])
    o.grid2particle(coords, index, pT, T)
    for it in range(25):
        o.particle2grid(coords, index, T, pT)
        o.advect(coords, index, 1, 2 / 3, V, dt)
        o.move(coords, index, [pT])
        o.inject(coords, index, [pT], 12, 42, it)
        o.grid2particle(coords, index, pT, T)
    assert abs(T.sum() - T0.sum()) / T0.sum() < 1e-2


def test_oracle_threads_match_serial():
    gr, o, coords, index = _setup(3, 8, True, S=24, nxcell=12)
    V = stream_velocity(gr)
    dt = cfl_dt(gr, V, 0.7)
    outs = []
    for nt in (1, 4):
        O.Oracle.set_threads(nt)
        c = [a.copy() for a in coords]; ix = index.copy(); pT = np.where(ix > 0, c[0], 0.0)
        o.advect(c, ix, 1, 0.5, V, dt)
        st = o.move(c, ix, [pT])
        inj = o.inject(c, ix, [pT], 10, 1, 0)
        outs.append((c, ix, pT, st, inj))
    O.Oracle.set_threads(1)
    a, b = outs
    assert a[3] == b[3] and a[4] == b[4] and np.array_equal(a[1], b[1])
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a[0], b[0]))
    assert np.array_equal(a[2], b[2], equal_nan=True)
