"""The opt-in "dense" slot policy of move_particles! (JP_MOVE_POLICY_DENSE in include/justpic_c.h; NOT reference behaviour).

Specification: (0) every live particle that fails the strict isincell test of its cell gives up its slot; (1) those leavers are
visited in the reference's order (3^N colours, source cells, slot order) and each one is deleted when it lies outside the domain,
else put into the LOWEST free slot of the cell the reference's bisection finds, else dropped (destination full).

CPU side of the evidence (the GPU side is tests/test_gpu_parity.py::test_move_policy_dense):
  * the oracle's twin (jpo_move_dense) == a second statement of the specification, written on the Julia transcription's building
    blocks (tests/test_move_inject_p2g_transcription.py: isincell / indomain / find_parent_cell_bisection, 1-based CellArrays),
    bit for bit on the adversarial states (faces, ulp gaps, far moves, NaN / Inf, full cells);
  * it keeps what a cell CONTAINS equal to the reference policy's whenever nothing is dropped, never drops more, and packs the
    cells lower (what the option is for);
  * a second call on a settled state changes nothing."""
import itertools
import math

import numpy as np
import pytest

from oracle import oracle as O
from tests.problems import make_grids, stream_velocity, cfl_dt
from tests.test_move_inject_p2g_transcription import (CASES, NAN, CellArray, adversarial_state, dxi_at, find_parent_cell_bisection, ids,
                                                      indomain, isincell, same)


def dense_move_particles(gr, co, idx, args):
    coords, index, fields = [CellArray(c) for c in co], CellArray(idx), [CellArray(a) for a in args]
    grid = [list(map(float, x)) for x in gr.xvi]
    nxi = index.size()
    domain_limits = [(min(x), max(x)) for x in grid]
    cells = [tuple(reversed(I)) for I in itertools.product(*[range(1, n + 1) for n in reversed(nxi)])]
    kept = {}                                                        # (ip, cell) -> payload of a leaver
    for I in cells:                                                  # (0) everybody vacates
        corner = tuple(grid[d][I[d] - 1] for d in range(len(nxi)))
        for ip in range(1, index.cellnum + 1):
            if index[(ip, *I)] == 0:
                continue
            p = tuple(c[(ip, *I)] for c in coords)
            if isincell(p, corner, dxi_at(gr, I)):
                continue
            kept[(ip, I)] = (p, tuple(a[(ip, *I)] for a in fields))
            index[(ip, *I)] = 0
            for a in coords + fields:
                a[(ip, *I)] = NAN
    counters = [0, 0, 0]                                             # moved, dropped, deleted
    n_color = [math.ceil(n / 3) for n in nxi]
    for offsets in itertools.product((1, 2, 3), repeat=len(nxi)):    # (1) place, in the reference's order
        for J in itertools.product(*[range(1, n + 1) for n in reversed(n_color)]):
            J = tuple(reversed(J))
            I = tuple(3 * (J[i] - 1) + offsets[i] for i in range(len(nxi)))
            if not all(I[i] <= nxi[i] for i in range(len(nxi))):
                continue
            for ip in range(1, index.cellnum + 1):
                if (ip, I) not in kept:
                    continue
                p, vals = kept[(ip, I)]
                if not indomain(p, domain_limits):
                    counters[2] += 1
                    continue
                new_cell = tuple(find_parent_cell_bisection(pi, x, s) for pi, x, s in zip(p, grid, I))
                free = next((i for i in range(1, index.cellnum + 1) if not index[(i, *new_cell)]), 0)
                if free == 0:
                    counters[1] += 1
                    continue
                index[(free, *new_cell)] = 1
                for c, v in zip(coords + fields, p + vals):
                    c[(free, *new_cell)] = v
                counters[0] += 1
    return tuple(counters)


@pytest.fixture(autouse=True)
def _reference_policy_afterwards():
    yield
    O.Oracle.set_move_policy("reference")


@pytest.mark.parametrize("case", CASES, ids=ids)
@pytest.mark.parametrize("seed", [1, 2])
def test_oracle_dense_twin_equals_the_specification(case, seed):
    ndim, n, uniform, S = case
    gr = make_grids(n, ndim, uniform, stretch=0.3)
    co, idx, args = adversarial_state(gr, S, np.random.default_rng(100 * seed + ndim))
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    O.Oracle.set_threads(1)
    co2, idx2, args2 = [c.copy() for c in co], idx.copy(), [a.copy() for a in args]
    for call in range(2):
        O.Oracle.set_move_policy("dense"); st = o.move(co, idx, args); O.Oracle.set_move_policy("reference")
        st2 = dense_move_particles(gr, co2, idx2, args2)
        assert st == st2, f"call {call}: (moved, dropped, deleted) {st} vs specification {st2}"
        same(idx, idx2, f"call {call}: index")
        for k, (a, b) in enumerate(zip(co + args, co2 + args2)):
            same(a, b, f"call {call}: array {k}")


def _top(idx):
    """1 + highest live slot per cell (0 for an empty cell)"""
    s = np.arange(1, idx.shape[0] + 1).reshape((-1,) + (1,) * (idx.ndim - 1))
    return (idx * s).max(axis=0)


@pytest.mark.parametrize("g", [(2, 24, True), (3, (9, 7, 12), False)], ids=lambda g: f"{g[0]}D-{g[1]}")
def test_dense_policy_same_cell_contents_lower_slots(g):
    ndim, n, uniform = g
    gr = make_grids(n, ndim, uniform)
    S, nx = 24, 12
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    co, idx = o.init_particles(nx, seed=7)
    V = stream_velocity(gr)
    dt = cfl_dt(gr, V, 0.9)
    T = np.where(idx > 0, np.nan_to_num(co[0]) * 3.0, np.nan)
    st8 = {"reference": None, "dense": None}
    state = {pol: ([c.copy() for c in co], idx.copy(), [T.copy()]) for pol in st8}
    drops = {pol: 0 for pol in st8}
    for it in range(8):
        for pol, (c, i, a) in state.items():
            o.advect(c, i, 1, 0.5, V, dt)
            O.Oracle.set_move_policy(pol); st = o.move(c, i, a); O.Oracle.set_move_policy("reference")
            drops[pol] += st[1]
        (cr, ir, ar), (cd, id_, ad) = state["reference"], state["dense"]
        assert drops["dense"] <= drops["reference"]
        if drops["reference"] == 0:
            for x, y in zip(cr + ar, cd + ad):
                assert np.array_equal(np.sort(np.nan_to_num(x, nan=np.inf), axis=0), np.sort(np.nan_to_num(y, nan=np.inf), axis=0)), \
                    f"step {it}: the two policies disagree on what a cell contains"
    (cr, ir, ar), (cd, id_, ad) = state["reference"], state["dense"]
    assert ir.sum() == id_.sum() or drops["reference"] > 0
    holes = lambda i: int((_top(i) - i.sum(axis=0)).sum())           # dead slots below the highest live one
    assert holes(id_) < 0.6 * holes(ir), (holes(id_), holes(ir))
    assert _top(id_).mean() < _top(ir).mean()
    # settled state: a second call moves nothing
    before = [x.copy() for x in cd + ad] + [id_.copy()]
    O.Oracle.set_move_policy("dense"); st = o.move(cd, id_, ad)
    assert st == (0, 0, 0)
    for x, y in zip(before, cd + ad + [id_]):
        same(x, y, "second dense call")
