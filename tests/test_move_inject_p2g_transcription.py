"""An INDEPENDENT pin of the oracle for move_particles!, inject_particles! and particle2grid! (SURVEY.md section 8c:
the reference holds no vectors for these three and Julia cannot run here).

oracle/justpic_oracle.c and the product's literal routines (csrc/jp_core.h) were restated by the same hand; a shared
misreading of the Julia would be invisible to GPU == oracle.  Here the three Julia routines are transcribed a second time,
line by line, in the reference's own terms -- 1-based indices, tuples, `CAI.@index A[ip, I...]` accessors, the same loop
nests and early exits, `muladd` / `fma` as exactly rounded fused operations -- and the C oracle must reproduce them bit for
bit on adversarial states: particles exactly on cell faces and vertices, in the fl(x + dx) ulp gap, moves of more than one
cell, particles leaving the domain, NaN / Inf coordinates, full destination cells (drops), the free-slot cursor carried
across destinations, empty cells and empty neighbourhoods.

    move_particles!    src/Particles/move_safe.jl:23-49 (colour loops), :51-62 (kernel), :72-125 (move_kernel!), :192-206
                       src/Particles/utils.jl:7-15 (isincell), src/Utils.jl:117-130 (find_parent_cell_bisection)
    inject_particles!  src/Particles/injection.jl:21-53, :55-66, :68-131, :330-393 (index_min_distance), :411-417, :442-461
    particle2grid!     src/Interpolations/particle_to_grid.jl:37-68 (2-D), :113-151 (3-D), :203-205, utils.jl:9-19 (distance)

Only the random numbers are ours (the reference calls the backend's rand): new_particle takes them from the library's
counter-based stream (oracle.rand3, keyed seed / step / cell / slot), which is pinned separately (tests/test_oracle_kat.py)."""
import itertools
import math
from fractions import Fraction

import numpy as np
import pytest

from oracle import oracle as O
from tests.problems import make_grids

NAN, INF = float("nan"), float("inf")


def fma(a, b, c):
    """muladd / fma on FMA hardware: one rounding"""
    if not (math.isfinite(a) and math.isfinite(b) and math.isfinite(c)):
        return a * b + c
    return float(Fraction(a) * Fraction(b) + Fraction(c))


class CellArray:
    """CAI.@index A[ip, i, j(, k)] on the (S, [nz,] ny, nx) arrays, 1-based as in Julia"""

    def __init__(self, a):
        self.a = a

    def __getitem__(self, key):
        ip, *I = key
        return self.a[(ip - 1, *[v - 1 for v in reversed(I)])].item()

    def __setitem__(self, key, val):
        ip, *I = key
        self.a[(ip - 1, *[v - 1 for v in reversed(I)])] = val

    @property
    def cellnum(self):
        return self.a.shape[0]

    def size(self):
        return tuple(reversed(self.a.shape[1:]))


def dxi_at(gr, idx):
    """@dxi di idx...: the scalar x[2] - x[1] of a range grid, diff(x)[I] of a vector grid"""
    return tuple(float(x[1] - x[0]) if gr.uniform else float(x[i] - x[i - 1]) for x, i in zip(gr.xvi, idx))


# ------------------------------------------------------------------------------------------------ move_particles!
def isincell(p, xci, dxi):
    b = True
    for pi, xv, dx in zip(p, xci, dxi):
        b = b & (xv < pi < xv + dx)
    return b


def indomain(p, domain_limits):
    for pi, (lo, hi) in zip(p, domain_limits):
        if not (lo < pi < hi):
            return False
    return True


def find_parent_cell_bisection(px, x, seed):
    lo, hi = 1, len(x)
    while True:
        if x[seed - 1] <= px <= x[seed]:
            return seed
        if x[seed - 1] < px:
            lo = seed
            seed = (hi + seed) // 2
        else:
            hi = seed
            seed = (lo + seed) // 2


def find_free_memory(initial_index, index, I):
    for i in range(initial_index, index.cellnum + 1):
        if not index[(i, *I)]:
            return i
    return 0


def move_kernel(coords, corner_xi, grid, dxi, index, domain_limits, args, idx, counters):
    starting_point = 1
    for ip in range(1, index.cellnum + 1):
        if index[(ip, *idx)] == 0:                                   # doskip
            continue
        p = tuple(c[(ip, *idx)] for c in coords)
        if isincell(p, corner_xi, dxi):
            continue
        if not indomain(p, domain_limits):
            index[(ip, *idx)] = 0
            for c in coords:
                c[(ip, *idx)] = NAN
            for a in args:
                a[(ip, *idx)] = NAN
            counters[2] += 1
            continue
        new_cell = tuple(find_parent_cell_bisection(pi, x, s) for pi, x, s in zip(p, grid, idx))
        current_args = tuple(a[(ip, *idx)] for a in args)
        index[(ip, *idx)] = 0
        for c in coords:
            c[(ip, *idx)] = NAN
        for a in args:
            a[(ip, *idx)] = NAN
        free_idx = find_free_memory(starting_point, index, new_cell)
        if free_idx == 0:
            counters[1] += 1
            continue
        starting_point = free_idx
        index[(free_idx, *new_cell)] = 1
        for c, v in zip(coords, p):
            c[(free_idx, *new_cell)] = v
        for a, v in zip(args, current_args):
            a[(free_idx, *new_cell)] = v
        counters[0] += 1


def julia_move_particles(gr, co, idx, args):
    coords, index, fields = [CellArray(c) for c in co], CellArray(idx), [CellArray(a) for a in args]
    grid = [list(map(float, x)) for x in gr.xvi]
    nxi = index.size()
    domain_limits = [(min(x), max(x)) for x in grid]
    n_color = [math.ceil(n / 3) for n in nxi]
    counters = [0, 0, 0]                                             # moved, dropped, deleted (diagnostics of the oracle)
    for offsets in itertools.product((1, 2, 3), repeat=len(nxi)):    # for offset_i in 1:3, offset_j in 1:3(, offset_k in 1:3)
        for I in itertools.product(*[range(1, n + 1) for n in reversed(n_color)]):
            I = tuple(reversed(I))                                   # the order of work-items within a launch is immaterial
            indices = tuple(3 * (I[i] - 1) + offsets[i] for i in range(len(nxi)))
            if all(indices[i] <= nxi[i] for i in range(len(nxi))):
                corner_xi = tuple(grid[d][indices[d] - 1] for d in range(len(nxi)))
                move_kernel(coords, corner_xi, grid, dxi_at(gr, indices), index, domain_limits, fields, indices, counters)
    return tuple(counters)


# ------------------------------------------------------------------------------------------------ inject_particles!
def distance(a, b):
    s = (a[0] - b[0]) ** 2
    for x, y in zip(a[1:], b[1:]):
        s = s + (x - y) ** 2
    return math.sqrt(s)


def index_min_distance(coords, pn, index, current_cell, cell):
    N = len(cell)
    particle_idx_min, cell_min = 0, (0,) * N
    dist_min = INF
    n = index.size()
    ranges = [range(c - 1, c + 2) for c in reversed(cell)]           # for k ..., j ..., i ..., ip in cellaxes(index)
    for rev in itertools.product(*ranges):
        I = tuple(reversed(rev))
        for ip in range(1, index.cellnum + 1):
            if any(v < 1 for v in I) or any(v > m for v, m in zip(I, n)):
                continue
            if I == tuple(cell) and ip == current_cell:
                continue
            if not index[(ip, *I)]:
                continue
            pxi = tuple(c[(ip, *I)] for c in coords)
            if N == 2 and any(math.isnan(v) for v in pxi):
                continue
            d = distance(pxi, pn)
            if d < dist_min:
                particle_idx_min, cell_min, dist_min = ip, I, d
    return particle_idx_min, cell_min


def quadrant_corners(xvi, dq):
    N = len(xvi)
    masks = [(0, 0), (1, 0), (0, 1), (1, 1)] if N == 2 else [(0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    return [xvi] + [tuple(x + d * m for x, d, m in zip(xvi, dq, mk)) for mk in masks[1:]]


def inject_cell(args, coords, index, grid, di_quadrant, min_xcell, idx_cell, rand, counters):
    xvi = tuple(grid[d][idx_cell[d] - 1] for d in range(len(idx_cell)))
    xvi_quadrants = quadrant_corners(xvi, di_quadrant)
    min_xQuadrant = -(-min_xcell // len(xvi_quadrants))              # cld
    for vertex in xvi_quadrants:
        particles_num = 0
        for i in range(1, index.cellnum + 1):
            if not index[(i, *idx_cell)]:
                continue
            pcoords = tuple(c[(i, *idx_cell)] for c in coords)
            if not isincell(pcoords, vertex, di_quadrant):
                continue
            particles_num += 1
        if particles_num >= min_xQuadrant:
            break
        for i in range(1, index.cellnum + 1):
            if index[(i, *idx_cell)]:
                continue
            particles_num += 1
            r = rand(idx_cell, i)
            p_new = tuple(x + d * fma(0.95, rr, 0.05) for x, d, rr in zip(vertex, di_quadrant, r))
            for c, v in zip(coords, p_new):
                c[(i, *idx_cell)] = v
            index[(i, *idx_cell)] = 1
            counters[0] += 1
            particle_idx, min_idx = index_min_distance(coords, p_new, index, i, idx_cell)
            for a in args:
                if particle_idx:                                     # (0, (0, 0)) when the neighbourhood holds no other particle: an
                    a[(i, *idx_cell)] = a[(particle_idx, *min_idx)]  # out-of-bounds read in the reference; the field is left as it was
            if particles_num >= min_xQuadrant:
                break


def julia_inject_particles(gr, co, idx, args, min_xcell, seed, step):
    coords, index, fields = [CellArray(c) for c in co], CellArray(idx), [CellArray(a) for a in args]
    grid = [list(map(float, x)) for x in gr.xvi]
    ni = index.size()
    N = len(ni)
    n_color = [math.ceil(n * 0.5) for n in ni]
    counters = [0]

    def rand(cell, slot):                                            # the library's counter-based stream, keyed by 0-based cell / slot
        c = 0
        for d in reversed(range(N)):
            c = c * ni[d] + (cell[d] - 1)
        return O.rand3(seed, 1, step, c, slot - 1)[:N]

    for offsets in itertools.product((1, 2), repeat=N):
        for I in itertools.product(*[range(1, n + 1) for n in n_color]):
            indices = tuple(2 * (I[i] - 1) + offsets[i] for i in range(N))
            if all(indices[i] <= ni[i] for i in range(N)):
                dq = tuple(d / 2 for d in dxi_at(gr, indices))
                inject_cell(fields, coords, index, grid, dq, min_xcell, indices, rand, counters)
    return counters[0]


# ------------------------------------------------------------------------------------------------ particle2grid!
def julia_particle2grid(gr, co, idx, Fp):
    coords, index, fp = [CellArray(c) for c in co], CellArray(idx), CellArray(Fp)
    N = gr.ndim
    sizeF = tuple(n + 1 for n in gr.n)
    F = np.zeros(tuple(reversed(sizeF)))
    xi = [list(map(float, x)) for x in gr.xvi]
    for node in itertools.product(*[range(1, s + 1) for s in sizeF]):
        xvertex = tuple(xi[d][node[d] - 1] for d in range(N))
        w, wF = 0.0, 0.0
        for rev in itertools.product((-1, 0), repeat=N):             # koffset, joffset, ioffset (outermost first)
            off = tuple(reversed(rev))
            vertex = tuple(node[d] + off[d] for d in range(N))
            if not all(1 <= vertex[d] < sizeF[d] for d in range(N)):
                continue
            for ip in range(1, index.cellnum + 1):
                if index[(ip, *vertex)] == 0:
                    continue
                p_i = tuple(c[(ip, *vertex)] for c in coords)
                with np.errstate(all="ignore"):                      # a particle exactly on the node: inv(0.0) = Inf, as in Julia
                    w_i = float(np.float64(1.0) / np.float64(distance(xvertex, p_i) ** 2))     # inv(distance(a, b)^order), order = 2
                w += w_i
                wF = fma(w_i, fp[(ip, *vertex)], wF)                 # 2-D: fma; 3-D: muladd
        with np.errstate(all="ignore"):
            val = (np.float64(wF) / np.float64(w)) if N == 2 else np.float64(wF) * (np.float64(1.0) / np.float64(w))
        F[tuple(v - 1 for v in reversed(node))] = val
    return F


# ------------------------------------------------------------------------------------------------ states
def adversarial_state(gr, S, rng, fill=0.55, far=0.05):
    """live particles anywhere within ~1.3 cells of their cell, some exactly on faces / vertices / in the ulp gap, some NaN / Inf /
    outside the domain, some cells full, some empty"""
    N, n = gr.ndim, gr.n
    shape = (S, *reversed(n))
    idx = (rng.random(shape) < fill).astype(np.uint8)
    cellsel = rng.random(tuple(reversed(n)))
    idx[:, cellsel < 0.08] = 1                                       # full cells: arrivals are dropped
    idx[:, cellsel > 0.94] = 0                                       # empty cells
    co = []
    for d in range(N):
        xv = np.asarray(gr.xvi[d], dtype=np.float64)
        ax = [1] * (N + 1); ax[N - d] = n[d]
        lo = xv[:-1].reshape(ax); dx = np.diff(xv).reshape(ax)
        u = rng.random(shape)
        kind = rng.random(shape)
        pos = lo + dx * u                                            # inside the cell
        pos = np.where(kind < 0.45, lo + dx * (u * 2.6 - 0.8), pos)  # up to ~1 cell away (some further)
        pos = np.where(kind < far, lo + dx * (u * 7.0 - 3.0), pos)   # far moves / out of the domain
        pos = np.where((kind > 0.90) & (kind < 0.93), lo, pos)       # exactly on the lower face
        pos = np.where((kind > 0.93) & (kind < 0.96), lo + dx, pos)  # on fl(x + dx): the upper face or its ulp gap
        pos = np.where((kind > 0.96) & (kind < 0.975), np.nextafter(lo + dx, -np.inf), pos)
        pos = np.where((kind > 0.975) & (kind < 0.985), np.broadcast_to(xv[np.minimum(np.arange(n[d]) + 2, n[d])].reshape(ax), shape), pos)
        co.append(pos)
    special = rng.random(shape)
    co[0] = np.where(special < 0.01, np.nan, co[0])
    co[N - 1] = np.where((special > 0.01) & (special < 0.02), np.inf, co[N - 1])
    for d in range(N):
        co[d] = np.where(idx > 0, co[d], np.nan)
    args = [np.where(idx > 0, rng.random(shape) * 10, np.nan), np.where(idx > 0, np.floor(rng.random(shape) * 3) + 1, np.nan)]
    return [np.ascontiguousarray(c) for c in co], np.ascontiguousarray(idx), [np.ascontiguousarray(a) for a in args]


CASES = [(2, (7, 5), True, 10), (2, (6, 8), False, 9), (3, (4, 5, 3), True, 8), (3, (5, 3, 4), False, 7)]
ids = lambda c: f"{c[0]}D-{c[1]}-{'range' if c[2] else 'vector'}-S{c[3]}"


def same(a, b, what):
    assert np.array_equal(a, b, equal_nan=True), f"{what}: {int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())} entries differ"


@pytest.mark.parametrize("case", CASES, ids=ids)
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_move_particles_oracle_equals_julia_transcription(case, seed):
    ndim, n, uniform, S = case
    gr = make_grids(n, ndim, uniform, stretch=0.3)
    rng = np.random.default_rng(100 * seed + ndim)
    co, idx, args = adversarial_state(gr, S, rng)
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    O.Oracle.set_threads(1)               # far moves race between same-colour cells (in the reference too): serial order on both sides
    co2, idx2, args2 = [c.copy() for c in co], idx.copy(), [a.copy() for a in args]
    for sweep in range(2):                # a second call re-slots what the first left on faces / in ulp gaps
        st = o.move(co, idx, args)
        st2 = julia_move_particles(gr, co2, idx2, args2)
        assert st == st2, f"call {sweep}: (moved, dropped, deleted) {st} vs transcription {st2}"
        same(idx, idx2, f"call {sweep}: index")
        for d in range(ndim):
            same(co[d], co2[d], f"call {sweep}: coords[{d}]")
        for k in range(2):
            same(args[k], args2[k], f"call {sweep}: args[{k}]")


def test_move_cases_cover_the_quirks():
    """the adversarial states do exercise what they are meant to: drops with free slots below the cursor, deletions, ties"""
    gr = make_grids((7, 5), 2, True)
    tot = np.zeros(3, dtype=np.int64)
    for seed in range(1, 4):
        co, idx, args = adversarial_state(gr, 10, np.random.default_rng(100 * seed + 2))
        tot += np.array(julia_move_particles(gr, co, idx, args))
    assert tot[0] > 50 and tot[1] > 5 and tot[2] > 5, tot


@pytest.mark.parametrize("case", CASES, ids=ids)
@pytest.mark.parametrize("seed", [4, 5])
def test_inject_particles_oracle_equals_julia_transcription(case, seed):
    ndim, n, uniform, S = case
    gr = make_grids(n, ndim, uniform, stretch=0.3)
    rng = np.random.default_rng(100 * seed + ndim)
    co, idx, args = adversarial_state(gr, S, rng, fill=0.3, far=0.0)
    # inject_particles! runs after move_particles!: every particle lies in its cell or on its faces; keep ties, drop the rest
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    O.Oracle.set_threads(1)
    o.move(co, idx, args)
    if seed == 5:                          # a corner region with no particle at all: the donor search finds nothing
        sl = (slice(None),) + (slice(0, 3),) * ndim
        idx[sl] = 0
        for a in co + args:
            a[sl] = np.nan
    co2, idx2, args2 = [c.copy() for c in co], idx.copy(), [a.copy() for a in args]
    min_xcell = 2 ** ndim * 2
    for step in range(2):
        inj = o.inject(co, idx, args, min_xcell, 77, step)
        inj2 = julia_inject_particles(gr, co2, idx2, args2, min_xcell, 77, step)
        assert inj == inj2 and (step > 0 or inj > 0)
        same(idx, idx2, f"step {step}: index")
        for d in range(ndim):
            same(co[d], co2[d], f"step {step}: coords[{d}]")
        for k in range(2):
            same(args[k], args2[k], f"step {step}: args[{k}]")


@pytest.mark.parametrize("case", CASES, ids=ids)
def test_particle2grid_oracle_equals_julia_transcription(case):
    ndim, n, uniform, S = case
    gr = make_grids(n, ndim, uniform, stretch=0.3)
    rng = np.random.default_rng(9 + ndim)
    co, idx, args = adversarial_state(gr, S, rng, fill=0.4, far=0.0)
    for d in range(ndim):                  # finite coordinates (a NaN coordinate poisons a node identically on both sides; keep a few)
        co[d] = np.where(np.isinf(co[d]), np.nan, co[d])
    o = O.Oracle(gr.xvi, gr.xci, gr.xi_vel, S, uniform)
    F = np.empty(tuple(reversed([v + 1 for v in n])))
    with np.errstate(all="ignore"):
        o.particle2grid(co, idx, F, args[0])
    F2 = julia_particle2grid(gr, co, idx, args[0])
    same(F, F2, "particle2grid")
    assert np.isnan(F2).any() and np.isfinite(F2).any()      # empty neighbourhoods give 0/0 = NaN as in the reference
