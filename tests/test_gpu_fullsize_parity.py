"""GPU == oracle, bit for bit, AT THE BASELINE.json CONFIGURATION SIZES.

The small-grid parity tests (tests/test_gpu_parity.py) prove every entry point on grids up to ~70 cells
wide; what they cannot reach is what only exists at size: 65 536-CTA grids, the paged destination-code
words, the 11 GB move workspace, TMA boxes at large extents, the advection -> move hand-off epilogue on
thousands of bricks.  Here the reference's four single-GPU configurations are run through the public API
(-> ctypes -> C ABI) and through the CPU oracle on the same seeded inputs, and compared after EVERY call:

  cfg1  scripts/temperature_advection.jl     2-D 256^2, 24 ppc (12 / 48), RK2, move, inject, g2p / p2g of T
  cfg2  scripts/rotating_circle.jl           2-D 512^2, RK4, inject_particles! reseeding, cell assignment
  cfg3  scripts/temperature_advection3D.jl   3-D 128^3, 24 ppc, RK2, trilinear g2p / p2g
  cfg4  BASELINE configs[3] (the headline)   3-D 256^3, 24 ppc, RK2 + hand-off, 3 advected fields, p2g, phase ratios

Bar: occupancy masks, slot assignment, coordinates, particle fields, exact-mode grid fields and centre phase
ratios bit-exact; the default (two-pass) particle2grid within the stated 1e-12.  The oracle runs with all
host threads (OpenMP over same-colour cells = the reference's own decomposition): ~2 s / step at 128^3,
~15 s / step at 256^3 on 16 threads; cfg4 needs ~45 GB of host memory and is skipped on smaller boxes.
cfg4 runs ONE coupled step by default (4 min, most of it the oracle and the 7 x 6.4 GB comparisons per call);
JP_CFG4_STEPS=2 adds a step that starts from fragmented slot planes (profiles/r02a_fullsize_parity_pytest.log: passed)."""
import os

import numpy as np
import pytest
import torch

from oracle.oracle import Oracle
from tests.problems import cfl_dt, make_grids, rotation_velocity, stream_velocity, vertex_field_linear

pytestmark = pytest.mark.gpu


def jp():
    import justpic.jl_b200 as J
    return J


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def same_cellarray(gpu, ref, what):
    """bit-level equality of a CellArray (NaN == NaN), compared on the device one slot plane at a time"""
    assert tuple(gpu.shape) == tuple(ref.shape), what
    bad = 0
    for s in range(ref.shape[0]):
        r = torch.from_numpy(ref[s]).cuda()
        g = gpu[s]
        if ref.dtype == np.uint8:
            bad += int((g != r).sum())
        else:
            bad += int((~((g == r) | (torch.isnan(g) & torch.isnan(r)))).sum())
    assert bad == 0, f"{what}: {bad} entries differ from the oracle"


def same_grid(gpu, ref, what, rtol=None):
    g = gpu.cpu().numpy()
    assert np.array_equal(np.isnan(g), np.isnan(ref)), f"{what}: NaN pattern"
    ok = ~np.isnan(ref)
    if rtol is None:
        nbad = int((g[ok] != ref[ok]).sum())
        assert nbad == 0, f"{what}: {nbad} entries differ from the oracle"
    else:
        scale = float(np.abs(ref[ok]).max())
        np.testing.assert_allclose(g[ok], ref[ok], rtol=rtol, atol=rtol * scale, err_msg=what)


def host_gb_free():
    try:
        import psutil
        return psutil.virtual_memory().available / 1e9
    except Exception:
        return 0.0


CFGS = {
    # name: ndim, n, scheme, velocity, CFL, steps, inject, fields
    "cfg1_2d_256_rk2": dict(ndim=2, n=256, rk4=False, rot=False, cfl=0.75, steps=5, inject=True, nfields=1, phases=0),
    "cfg2_2d_512_rk4_inject": dict(ndim=2, n=512, rk4=True, rot=True, cfl=None, steps=5, inject=True, nfields=1, phases=0),
    "cfg3_3d_128_rk2": dict(ndim=3, n=128, rk4=False, rot=False, cfl=0.5, steps=3, inject=True, nfields=1, phases=0),
    "cfg4_3d_256_headline": dict(ndim=3, n=256, rk4=False, rot=False, cfl=0.5, steps=int(os.environ.get("JP_CFG4_STEPS", "1")),
                                 inject=False, nfields=3, phases=2),
}


@pytest.mark.parametrize("name", list(CFGS))
def test_baseline_config_bit_exact_vs_oracle(name):
    J = jp()
    cfg = CFGS[name]
    ndim, n = cfg["ndim"], cfg["n"]
    big = name.startswith("cfg4")
    if big and (torch.cuda.mem_get_info()[1] < 100e9 or host_gb_free() < 60):
        pytest.skip("cfg4 needs ~60 GB of device memory and ~45 GB of host memory")
    gr = make_grids(n, ndim, True)
    p = J.init_particles(J.CUDABackend, 24, 48, 12, *gr.grid_vel, seed=42)
    o = Oracle(gr.xvi, gr.xci, gr.xi_vel, p.max_xcell, True)
    co, idx = o.init_particles(24, 42)

    def check(what, gargs=(), oargs=()):
        for d in range(ndim):
            same_cellarray(p.coords[d], co[d], f"{name} {what}: coords[{d}]")
        same_cellarray(p.index, idx, f"{name} {what}: index")
        for i, (a, b) in enumerate(zip(gargs, oargs)):
            same_cellarray(a, b, f"{name} {what}: args[{i}]")

    check("init_particles")
    if cfg["rot"]:
        # scripts/rotating_circle.jl:19,27,57: solid rotation, dt = 200 on 200 cells (0.63 cells per step at the rim);
        # the same Courant number on n cells -- beyond one cell per step the reference's colour sweeps race
        V = rotation_velocity(gr); dt = 200.0 * 200 / n
    else:
        V = stream_velocity(gr); dt = cfl_dt(gr, V, cfg["cfl"])
    Vd = [dev(v) for v in V]
    method, scheme, alpha = (J.RungeKutta4(), 2, 0.0) if cfg["rk4"] else (J.RungeKutta2(), 1, 0.5)
    T = vertex_field_linear(gr); Td = dev(T)
    gargs = list(J.init_cell_arrays(p, cfg["nfields"]))
    oargs = [np.zeros_like(co[0]) for _ in range(cfg["nfields"])]
    J.grid2particle(gargs[0], Td, p); o.grid2particle(co, idx, oargs[0], T)
    same_cellarray(gargs[0], oargs[0], f"{name} grid2particle")
    if cfg["rot"]:                                   # the circle's phase field rides along as the particle field
        r2 = (co[0] - 0.5) ** 2 + (co[1] - 0.75) ** 2
        oargs[0][...] = np.where(idx > 0, 1.0 + (r2 < 0.15 ** 2), 0.0)
        gargs[0].copy_(dev(oargs[0]))
    if cfg["nfields"] == 3:                          # (T, phase, strain)
        oargs[1][...] = np.where(idx > 0, 1.0 + (co[0] < co[2]), 0.0)
        oargs[2][...] = np.where(idx > 0, co[1] * co[0], 0.0)
        gargs[1].copy_(dev(oargs[1])); gargs[2].copy_(dev(oargs[2]))
    K = cfg["phases"]
    pr = J.PhaseRatios(J.CUDABackend, K, gr.n) if K else None
    Tg = torch.empty_like(Td); oT = np.empty_like(T)
    moved = injected = 0
    for it in range(cfg["steps"]):
        J.advection(p, method, Vd, dt, classify=big)       # the headline configuration runs with the hand-off, as bench.py does
        o.advect(co, idx, scheme, alpha, V, dt)
        check(f"step {it} advection")
        J.move_particles(p, gargs); st = o.move(co, idx, oargs)
        assert J.last_move_path(p) == "plan", "the full-size run must exercise the plan / gather / scatter path"
        if big:
            assert J.last_move_classify(p) == "handoff"
        check(f"step {it} move_particles", gargs, oargs)
        assert J.move_stats(p) == st
        moved += st[0]
        if cfg["inject"]:
            J.inject_particles(p, gargs, step=it); inj = o.inject(co, idx, oargs, 12, 42, it)
            check(f"step {it} inject_particles", gargs, oargs)
            assert J.inject_stats(p) == inj
            injected += inj
        o.particle2grid(co, idx, oT, oargs[0])
        J.particle2grid(Tg, gargs[0], p, mode="exact")
        same_grid(Tg, oT, f"{name} step {it} particle2grid[exact]")
        J.particle2grid(Tg, gargs[0], p)                 # default mode: deterministic two-pass, stated tolerance
        same_grid(Tg, oT, f"{name} step {it} particle2grid[default]", rtol=1e-12)
        if K:
            ratios = np.zeros(o.cell_shape(K))
            J.phase_ratios_center(pr, p, gargs[1]); o.phase_ratios_center(co, ratios, oargs[1], K)
            same_cellarray(pr.center, ratios, f"{name} step {it} phase_ratios_center")
        if not big:                                      # the reference loop re-interpolates T each step (test/test_2D.jl:517-524)
            if not cfg["rot"]:
                J.grid2particle(gargs[0], Tg, p); o.grid2particle(co, idx, oargs[0], Tg.cpu().numpy())
                same_cellarray(gargs[0], oargs[0], f"{name} step {it} grid2particle")
    assert moved > 0
    if cfg["rot"]:
        assert injected >= 0
