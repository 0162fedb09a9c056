"""GPU parity of the advection -> move hand-off (JP_OPT_ADVECT_CLASSIFY, include/justpic_c.h):
`advection(..., classify=True)` leaves one classification byte per slot and the next
`move_particles` plans from those bytes instead of re-reading the coordinates.  The results must
be those of the reference's `move_particles!` (src/Particles/move_safe.jl:21-125) bit for bit --
checked against the oracle exactly like the default path -- and the bytes must be dropped
whenever the particles change in between."""
import numpy as np
import pytest
import torch

from tests.problems import cfl_dt, stream_velocity, vertex_field_linear
from tests.test_gpu_parity import GRIDS, Twin, dev, host, ids, jp

pytestmark = pytest.mark.gpu


def _fields(J, t):
    T = vertex_field_linear(t.gr)
    pT, ph = J.init_cell_arrays(t.p, 2)
    J.grid2particle(pT, dev(T), t.p)
    opT = np.zeros_like(t.co[0]); t.o.grid2particle(t.co, t.idx, opT, T)
    oph = np.where(t.idx > 0, 1.0 + (t.co[0] < t.co[-1]), 0.0)
    ph.copy_(dev(oph))
    return pT, ph, opT, oph


@pytest.mark.parametrize("g", GRIDS + [(3, (40, 9, 6), True), (2, (70, 11), True)], ids=ids)
def test_handoff_trajectory(g):
    J = jp()
    t = Twin(*g)
    V = stream_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(t.gr, V, 0.9)
    pT, ph, opT, oph = _fields(J, t)
    methods = [(J.RungeKutta2(), 1, 0.5), (J.RungeKutta4(), 2, 0.0), (J.RungeKutta2(2 / 3), 1, 2 / 3), (J.Euler(), 0, 0.0)]
    for it in range(8):
        m = methods[it % 4]
        J.advection(t.p, m[0], Vd, dt, classify=True); t.o.advect(t.co, t.idx, m[1], m[2], V, dt)
        t.check_state(f"step {it} advection (hand-off on)")
        J.move_particles(t.p, (pT, ph)); st = t.o.move(t.co, t.idx, [opT, oph])
        t.check_state(f"step {it} move_particles from hand-off bytes", (pT, ph), (opT, oph))
        assert J.move_stats(t.p) == st
        assert J.last_move_path(t.p) == "plan" and J.last_move_classify(t.p) == "handoff"
        J.inject_particles(t.p, (pT, ph), step=it); t.o.inject(t.co, t.idx, [opT, oph], t.min_xcell, t.seed, it)
        t.check_state(f"step {it} inject_particles", (pT, ph), (opT, oph))


@pytest.mark.parametrize("ndim", [2, 3])
def test_handoff_is_dropped_when_particles_change(ndim):
    """inject / clean / a second advect / other arrays between the two calls: the bytes must not be used."""
    J = jp()
    t = Twin(ndim, 12 if ndim == 2 else 8, True)
    V = stream_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(t.gr, V, 0.7)
    pT, ph, opT, oph = _fields(J, t)
    # (a) inject in between
    J.advection(t.p, J.RungeKutta2(), Vd, dt, classify=True); t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
    J.inject_particles(t.p, (pT, ph), step=0); t.o.inject(t.co, t.idx, [opT, oph], t.min_xcell, t.seed, 0)
    J.move_particles(t.p, (pT, ph)); t.o.move(t.co, t.idx, [opT, oph])
    assert J.last_move_classify(t.p) == "coords"
    t.check_state("inject between advect and move", (pT, ph), (opT, oph))
    # (b) two advects in a row: the bytes of the second one are the valid ones
    J.advection(t.p, J.RungeKutta2(), Vd, dt); t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
    J.advection(t.p, J.Euler(), Vd, 0.3 * dt); t.o.advect(t.co, t.idx, 0, 0.0, V, 0.3 * dt)
    J.move_particles(t.p, (pT, ph)); t.o.move(t.co, t.idx, [opT, oph])
    t.check_state("two advects then move", (pT, ph), (opT, oph))
    # (c) a move consumes the bytes: a second move right after classifies the coordinates
    J.advection(t.p, J.RungeKutta2(), Vd, dt); t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
    J.move_particles(t.p, (pT, ph)); t.o.move(t.co, t.idx, [opT, oph])
    assert J.last_move_classify(t.p) == "handoff"
    J.move_particles(t.p, (pT, ph)); t.o.move(t.co, t.idx, [opT, oph])
    assert J.last_move_classify(t.p) == "coords"
    t.check_state("second move", (pT, ph), (opT, oph))
    # (d) clean in between
    J.advection(t.p, J.RungeKutta2(), Vd, dt); t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
    J.clean_particles(t.p, None, (pT, ph)); t.o.clean(t.co, t.idx, [opT, oph])
    J.move_particles(t.p, (pT, ph)); t.o.move(t.co, t.idx, [opT, oph])
    assert J.last_move_classify(t.p) == "coords"
    t.check_state("clean between advect and move", (pT, ph), (opT, oph))
    # (e) switched off again
    J.advection(t.p, J.RungeKutta2(), Vd, dt, classify=False); t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
    J.move_particles(t.p, (pT, ph)); t.o.move(t.co, t.idx, [opT, oph])
    assert J.last_move_classify(t.p) == "coords"
    t.check_state("hand-off off", (pT, ph), (opT, oph))


@pytest.mark.parametrize("ndim", [2, 3])
def test_handoff_with_halo_unpack_in_between(ndim):
    """The multi-GPU time loop: advection! -> update_halo! -> move_particles!.  Planes rewritten by
    jp_halo_unpack are re-classified from the coordinates, the rest comes from the words advect left.
    The boundary planes are overwritten (through jp_halo_pack / jp_halo_unpack) with what they held BEFORE
    the advection: every particle there is inside its cell again, while the hand-off says ~40 % of them
    left -- a move that trusted the stale words would relocate particles that must stay.  (Planes holding
    a neighbour's particles would do as well, but particles up to two cells from their halo cell take the
    literal sweeps, which only the serial oracle makes deterministic.)"""
    J = jp()
    from justpic.jl_b200 import halo as H
    n = 12 if ndim == 2 else (8, 6, 7)
    t = Twin(ndim, n, True)
    gr = t.gr
    V = stream_velocity(gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(gr, V, 0.9)
    pT, ph, opT, oph = _fields(J, t)
    for it in range(3):
        arrays = [*t.p.coords, pT, ph]
        oarrays = [*t.co, opT, oph, t.idx]
        planes = [(d, pl) for d in range(ndim) for pl in (0, gr.n[d] - 1)]
        bufs, saved = [], []
        for d, pl in planes:
            buf = torch.empty(H.plane_bytes(t.p.ncells, t.p.max_xcell, d, len(arrays)), dtype=torch.uint8, device="cuda")
            H._cuda_pack(t.p, d, pl, arrays, buf)
            bufs.append(buf)
            sl = [slice(None)] * oarrays[0].ndim; sl[ndim - d] = pl          # arrays are (S, [nz,] ny, nx)
            saved.append([a[tuple(sl)].copy() for a in oarrays])
        J.advection(t.p, J.RungeKutta2(), Vd, dt, classify=True); t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
        for (d, pl), buf, sv in zip(planes, bufs, saved):
            H._cuda_unpack(t.p, d, pl, arrays, buf)
            sl = [slice(None)] * oarrays[0].ndim; sl[ndim - d] = pl
            for a, v in zip(oarrays, sv):
                a[tuple(sl)] = v
        t.check_state(f"step {it} after restoring the boundary planes", (pT, ph), (opT, oph))
        J.move_particles(t.p, (pT, ph)); st = t.o.move(t.co, t.idx, [opT, oph])
        assert J.last_move_classify(t.p) == "handoff" and J.last_move_path(t.p) == "plan"
        t.check_state(f"step {it} move after halo unpack", (pT, ph), (opT, oph))
        assert J.move_stats(t.p) == st
        J.inject_particles(t.p, (pT, ph), step=it); t.o.inject(t.co, t.idx, [opT, oph], t.min_xcell, t.seed, it)


@pytest.mark.parametrize("ndim", [2, 3])
def test_handoff_ties_fall_back_to_direct_sweeps(ndim):
    """Particles that end an advection exactly on a face / outside the domain: the bytes carry the
    'cannot plan' reason, the call takes the literal sweeps, results stay exact."""
    J = jp()
    t = Twin(ndim, 8, True, exact=True)
    gr = t.gr
    # zero velocity: positions stay where we put them -- on vertices, centres, outside
    V = [np.zeros_like(v) for v in stream_velocity(gr)]; Vd = [dev(v) for v in V]
    rng = np.random.default_rng(3)
    live = np.argwhere(t.idx > 0)
    for tt in live[rng.choice(len(live), size=len(live) // 4, replace=False)]:
        for d in range(ndim):
            if rng.random() < 0.5:
                cell = tt[ndim - d]
                cand = [gr.xvi[d][cell], gr.xvi[d][cell + 1], gr.xvi[d][min(cell + 2, gr.n[d])] - 1e-9, gr.xvi[d][max(cell - 1, 0)] + 1e-9]
                t.co[d][tuple(tt)] = cand[rng.integers(len(cand))]
    for d in range(ndim):
        t.p.coords[d].copy_(dev(t.co[d]))
    pT, = J.init_cell_arrays(t.p, 1)
    opT = np.where(t.idx > 0, rng.random(t.idx.shape), 0.0); pT.copy_(dev(opT))
    J.advection(t.p, J.RungeKutta2(), Vd, 1e-3, classify=True); t.o.advect(t.co, t.idx, 1, 0.5, V, 1e-3)
    t.check_state("advection with zero velocity")
    J.move_particles(t.p, (pT,)); st = t.o.move(t.co, t.idx, [opT])
    assert J.last_move_classify(t.p) == "handoff" and J.last_move_path(t.p) == "direct"
    t.check_state("move from tie positions", (pT,), (opT,))
    assert J.move_stats(t.p) == st


@pytest.mark.parametrize("ndim", [2, 3])
def test_invalidate_handoffs_after_a_foreign_write(ndim):
    """A write to the coordinates that does not go through the library (here: a host-side shift of some particles by a third of a cell)
    between advection! and move_particles! is invisible to the hand-off; jp_invalidate_handoffs tells the library, and
    move_particles! then classifies from the arrays as without the option."""
    J = jp()
    t = Twin(ndim, 14 if ndim == 2 else 9, True)
    V = stream_velocity(t.gr); Vd = [dev(v) for v in V]
    dt = cfl_dt(t.gr, V, 0.6)
    pT, ph, opT, oph = _fields(J, t)
    for it in range(3):
        J.advection(t.p, J.RungeKutta2(), Vd, dt, classify=True); t.o.advect(t.co, t.idx, 1, 0.5, V, dt)
        live = t.idx > 0
        sel = live & (np.arange(t.idx.size).reshape(t.idx.shape) % 11 == it)
        dx = float(t.gr.xvi[0][1] - t.gr.xvi[0][0])
        t.co[0][sel] += 0.35 * dx                        # the "boundary fix-up" of some host code (total displacement stays below one cell)
        t.p.coords[0].copy_(dev(t.co[0]))
        J.invalidate_handoffs(t.p)
        J.move_particles(t.p, (pT, ph)); st = t.o.move(t.co, t.idx, [opT, oph])
        assert J.last_move_classify(t.p) == "coords"
        t.check_state(f"step {it} move after a foreign write", (pT, ph), (opT, oph))
        assert J.move_stats(t.p) == st
