"""The move plan as an ALGORITHM, without a GPU.

`move_particles!` in the library is not the reference's 3^N read-modify-write sweeps over the particle arrays but
classify -> plan (3^N colour passes over per-cell occupancy WORDS) -> finalize / scan -> gather -> scatter
(justpic/jl_b200/csrc/jp_move_plan.cuh).  The GPU tests check the CUDA code against the oracle; this test checks
the decomposition itself: a plain-Python restatement of what each kernel computes, driven by the product's own classification
routine (jp_classify_particle, host build of jp_core.h), must give the oracle's slot assignment, payloads and counters bit for
bit, for every case the planner accepts (<= 1 cell per step, no ties); the other cases must raise the hand-over flag."""
import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.emul.emul import Emul
from tests.problems import cfl_dt, make_grids, stream_velocity

STAY, DELETE = 28, 27


def planned_move(e, gr, S, coords, index, args, policy="reference"):
    """Returns (moved, dropped, deleted) or None when the planner hands over to the direct sweeps.
    policy: "reference" (carried-over cursor), "compact" (every search from slot 0), "dense" (k_move_prevacate first: all leavers give
    up their slots before the first migrant is placed; the plan then neither clears bits nor carries a cursor)."""
    ndim = gr.ndim
    n = list(gr.n) + [1] * (3 - ndim)
    C = int(np.prod(n))
    flat = lambda a: a.reshape(S, C)
    co = [flat(c) for c in coords]; idx = flat(index); ar = [flat(a) for a in args]
    cell = lambda i, j, k: i + n[0] * (j + n[1] * k)
    # ---- A. k_move_classify3: occupancy word, leave word, destination codes of the leavers in slot order
    occ = [0] * C; leave = [0] * C; codes = [[] for _ in range(C)]
    for c in range(C):
        ci = (c % n[0], (c // n[0]) % n[1], c // (n[0] * n[1]))
        for s in range(S):
            if not idx[s, c]:
                continue
            occ[c] |= 1 << s
            code = e.classify(ci[:ndim], [float(co[d][s, c]) for d in range(ndim)])[1]
            if code == STAY:
                continue
            if code > STAY:
                return None                                  # complex flag: ties / far moves -> direct sweeps
            leave[c] |= 1 << s
            codes[c].append(code)
    occ0 = list(occ)
    if policy == "dense":                                    # k_move_prevacate
        occ = [occ0[c] & ~leave[c] for c in range(C)]
    # ---- B. k_move_plan x 3^N in the reference's colour order (offset_i outermost), words only
    res = [[None] * len(codes[c]) for c in range(C)]
    dropped = deleted = 0
    smask = (1 << S) - 1
    for ox in range(3):
        for oy in range(3):
            for oz in range(3 if ndim == 3 else 1):
                for k in range(oz, n[2], 3):
                    for j in range(oy, n[1], 3):
                        for i in range(ox, n[0], 3):
                            c = cell(i, j, k)
                            cursor = 0; occ_c = occ[c]; kk = 0
                            for ip in range(S):
                                if not (leave[c] >> ip) & 1:
                                    continue
                                code = codes[c][kk]; kk += 1
                                if policy != "dense":
                                    occ_c &= ~(1 << ip)
                                if code == DELETE:
                                    deleted += 1
                                    continue
                                dv = (code % 3 - 1, (code // 3) % 3 - 1, code // 9 - 1)
                                c2 = cell(i + dv[0], j + dv[1], k + dv[2])
                                free = ~occ[c2] & smask & ~((1 << cursor) - 1)
                                if free == 0:
                                    dropped += 1
                                    continue
                                fs = (free & -free).bit_length() - 1
                                if policy == "reference":
                                    cursor = fs               # the reference's carried-over starting_point
                                occ[c2] |= 1 << fs
                                res[c][kk - 1] = (c2, fs)
                            if policy != "dense":
                                occ[c] = occ_c
    # ---- C. k_move_finalize + exclusive scan
    arr = [occ[c] & (~occ0[c] | leave[c]) & smask for c in range(C)]
    off = np.concatenate([[0], np.cumsum([bin(a).count("1") for a in arr])]).astype(np.int64)
    # ---- D. k_move_gather: payload of every placed leaver -> staging[off[dest] + rank of its slot among the arrivals]
    arrays = co + ar
    stage = np.full((int(off[-1]), len(arrays)), np.nan)
    for c in range(C):
        kk = 0
        for ip in range(S):
            if not (leave[c] >> ip) & 1:
                continue
            r = res[c][kk]; kk += 1
            if r is None:
                continue
            c2, fs = r
            pos = off[c2] + bin(arr[c2] & ((1 << fs) - 1)).count("1")
            stage[pos] = [a[ip, c] for a in arrays]
    # ---- E. k_move_scatter: arrivals from staging, NaN into vacated slots that stayed empty, mask bytes
    for c in range(C):
        for s in range(S):
            a_bit, l_bit = (arr[c] >> s) & 1, (leave[c] >> s) & 1
            if not (a_bit or l_bit):
                continue
            if a_bit:
                pos = off[c] + bin(arr[c] & ((1 << s) - 1)).count("1")
                for q, a in enumerate(arrays):
                    a[s, c] = stage[pos, q]
                idx[s, c] = 1
            else:
                for a in arrays:
                    a[s, c] = np.nan
                idx[s, c] = 0
    return int(off[-1]), dropped, deleted


CASES = [
    # ndim, n, uniform, S, nxcell, cfl
    (2, (13, 10), True, 24, 12, 0.9),
    (2, (11, 9), False, 20, 12, 0.8),
    (3, (7, 5, 6), True, 24, 12, 0.9),
    (3, (5, 6, 4), False, 20, 10, 0.7),
    (2, (14, 11), True, 12, 12, 0.95),          # tight storage: arrivals are dropped
    (3, (6, 5, 5), True, 10, 8, 0.95),
]


@pytest.mark.parametrize("policy", ["reference", "compact", "dense"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}D-{c[1]}-{'range' if c[2] else 'vector'}-S{c[3]}")
def test_plan_gather_scatter_equals_literal_sweeps(case, policy):
    """(the oracle runs the same slot policy: its reference sweeps, or the twins of the library's two opt-in rules)"""
    ndim, n, uniform, S, nxcell, cfl = case
    gr = make_grids(n, ndim, uniform=uniform, stretch=0.4)
    o = Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    e = Emul(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    co, idx = o.init_particles(nxcell, 9)
    V = stream_velocity(gr); dt = cfl_dt(gr, V, cfl)
    f1 = np.where(idx > 0, co[0] * 3.0 + 1.0, 0.0); f2 = np.where(idx > 0, co[-1] - 0.5, 0.0)
    planned = handed_over = total_dropped = 0
    for it in range(5):
        o.advect(co, idx, 1, 0.5, V, dt)
        A = [[c.copy() for c in co], idx.copy(), [f1.copy(), f2.copy()]]
        st = planned_move(e, gr, S, A[0], A[1], A[2], policy)
        Oracle.set_move_policy(policy)
        try:
            ref = o.move(co, idx, [f1, f2])
        finally:
            Oracle.set_move_policy("reference")
        if st is None:
            handed_over += 1
            continue
        planned += 1; total_dropped += st[1]
        assert st == ref, f"step {it}: counters {st} vs oracle {ref}"
        assert np.array_equal(A[1], idx), f"step {it}: occupancy masks"
        for d in range(ndim):
            assert np.array_equal(A[0][d], co[d], equal_nan=True), f"step {it}: coords[{d}]"
        assert np.array_equal(A[2][0], f1, equal_nan=True) and np.array_equal(A[2][1], f2, equal_nan=True), f"step {it}: fields"
    assert planned >= 4, (planned, handed_over)
    if S <= 12:
        assert total_dropped > 0                         # the tight cases exercise the drop branch (under every policy)
