"""grid2particle_flip! and subgrid_diffusion! / subgrid_diffusion_centroid!.
CPU: the reference's own property for the flip kernel (test/test_interpolation_kernels.jl:93-99,
test/test_3D.jl:88-95: with T == T0 linear, pT stays ≈ the coordinate) and sanity of the subgrid
composition on the oracle.  GPU: both against the oracle -- flip bit for bit, subgrid diffusion
within the stated 1e-12 (it evaluates exp())."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from tests.problems import centre_field_linear, make_grids, vertex_field_linear

CASES = [(2, (12, 9), True), (2, (9, 14), False), (3, (6, 5, 7), True), (3, (5, 6, 4), False)]
cid = lambda c: f"{c[0]}D-{'x'.join(map(str, c[1]))}-{'range' if c[2] else 'vector'}"


def _setup(ndim, n, uniform, S=16, nxcell=8, seed=4):
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.3)
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    co, idx = o.init_particles(nxcell, seed)
    return gr, o, co, idx


@pytest.mark.parametrize("case", [c for c in CASES if c[2]], ids=cid)
def test_oracle_flip_linear_field_property(case):
    gr, o, co, idx = _setup(*case)
    T = vertex_field_linear(gr)
    pT = np.zeros_like(co[0])
    o.grid2particle(co, idx, pT, T)
    live = idx > 0
    np.testing.assert_allclose(pT[live], co[-1][live], rtol=math.sqrt(np.finfo(float).eps))
    o.grid2particle_flip(co, idx, pT, T, T.copy(), 0.0)
    np.testing.assert_allclose(pT[live], co[-1][live], rtol=math.sqrt(np.finfo(float).eps))
    # alpha = 1: pure PIC = grid2particle with the grid_size spacing
    pT2 = np.full_like(pT, 123.0)
    o.grid2particle_flip(co, idx, pT2, T, 0 * T, 1.0)
    np.testing.assert_allclose(pT2[live], co[-1][live], rtol=1e-12)
    assert np.all(pT2[~live] == 123.0)


def _subgrid_inputs(gr, co, idx, centroid, rng):
    N = gr.ndim
    Tg = (centre_field_linear(gr) if centroid else vertex_field_linear(gr)) ** 2 + 1.0
    ext = [v + (1 if centroid else 2) for v in gr.n]                     # dT_grid: one ghost node on the low side (read at I + 1)
    dT = np.ascontiguousarray(0.01 * rng.standard_normal(tuple(reversed(ext))))
    pT = np.where(idx > 0, 1.0 + co[-1] ** 2 + 0.05 * rng.standard_normal(idx.shape), np.nan)
    dt0 = np.where(idx > 0, rng.uniform(0.2, 2.0, idx.shape), np.nan)
    return np.ascontiguousarray(Tg), dT, pT, dt0


@pytest.mark.parametrize("case", CASES, ids=cid)
@pytest.mark.parametrize("centroid", [False, True], ids=["vertex", "centroid"])
def test_oracle_subgrid_diffusion_sanity(case, centroid):
    gr, o, co, idx = _setup(*case)
    rng = np.random.default_rng(0)
    Tg, dT, pT, dt0 = _subgrid_inputs(gr, co, idx, centroid, rng)
    pT_in = pT.copy()
    pT0 = np.zeros_like(pT); pdT = np.zeros_like(pT)
    sub = np.zeros(Tg.shape)
    o.subgrid_diffusion(co, idx, pT, Tg, dT, pT0, pdT, dt0, sub, 0.5, 1.0, centroid)
    live = idx > 0
    assert np.all(np.isfinite(pT[live])) and np.all(np.isfinite(pT0[live]))
    # d = 0: no relaxation -> pT0 stays the old temperature and pdT = interpolated(dT - p2g(0)) = interpolated dT
    pT2 = pT_in.copy(); pT0b = np.zeros_like(pT); pdTb = np.zeros_like(pT); sub2 = np.zeros(Tg.shape)
    o.subgrid_diffusion(co, idx, pT2, Tg, dT, pT0b, pdTb, dt0, sub2, 0.5, 0.0, centroid)
    np.testing.assert_array_equal(pT0b[live], pT_in[live])
    np.testing.assert_allclose(pT2[live], pT_in[live] + pdTb[live], rtol=0, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=cid)
def test_gpu_grid2particle_flip(case):
    import torch
    import justpic.jl_b200 as J
    gr, o, co, idx = _setup(*case)
    grids = gr.grid_vel if gr.uniform else gr.xi_vel
    p = J.init_particles(J.CUDABackend, 8, 16, 4, *grids, seed=4)
    rng = np.random.default_rng(3)
    T = np.ascontiguousarray(vertex_field_linear(gr) + 0.2 * rng.standard_normal(vertex_field_linear(gr).shape))
    T0 = np.ascontiguousarray(T + 0.1 * rng.standard_normal(T.shape))
    opT = np.where(idx > 0, rng.standard_normal(idx.shape), 7.0)
    for alpha in (0.0, 0.3, 1.0):
        pT = torch.from_numpy(opT).cuda()
        J.grid2particle_flip(pT, None, torch.from_numpy(T).cuda(), torch.from_numpy(T0).cuda(), p, alpha=alpha)
        ref = opT.copy()
        o.grid2particle_flip(co, idx, ref, T, T0, alpha)
        assert np.array_equal(pT.cpu().numpy(), ref), f"alpha {alpha}: {(pT.cpu().numpy() != ref).sum()} entries differ"
        opT = ref


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=cid)
@pytest.mark.parametrize("centroid", [False, True], ids=["vertex", "centroid"])
def test_gpu_subgrid_diffusion(case, centroid):
    import torch
    import justpic.jl_b200 as J
    gr, o, co, idx = _setup(*case)
    grids = gr.grid_vel if gr.uniform else gr.xi_vel
    p = J.init_particles(J.CUDABackend, 8, 16, 4, *grids, seed=4)
    rng = np.random.default_rng(1)
    Tg, dT, opT, odt0 = _subgrid_inputs(gr, co, idx, centroid, rng)
    sa = J.SubgridDiffusionCellArrays(p, loc="center" if centroid else "vertex")
    sa.dt0.copy_(torch.from_numpy(odt0))
    pT = torch.from_numpy(opT).cuda()
    opT0 = np.zeros_like(opT); opdT = np.zeros_like(opT); osub = np.zeros(Tg.shape)
    J.api.P2G_MODE = "exact"
    try:
        fn = J.subgrid_diffusion_centroid if centroid else J.subgrid_diffusion
        for it in range(2):
            fn(pT, torch.from_numpy(Tg).cuda(), torch.from_numpy(dT).cuda(), sa, p, 0.37, d=0.8)
            o.subgrid_diffusion(co, idx, opT, Tg, dT, opT0, opdT, odt0, osub, 0.37, 0.8, centroid)
    finally:
        J.api.P2G_MODE = "twopass_fastw"
    live = idx > 0

    def close(a, b, what):
        a = a.cpu().numpy()
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern"
        ok = ~np.isnan(b)
        scale = np.abs(b[ok]).max()
        np.testing.assert_allclose(a[ok], b[ok], rtol=1e-12, atol=1e-12 * scale, err_msg=what)

    close(pT, opT, "pT"); close(sa.pT0, opT0, "pT0"); close(sa.pdT, opdT, "pdT"); close(sa.dT_subgrid, osub, "dT_subgrid")
    assert np.all(np.isfinite(pT.cpu().numpy()[live]))
