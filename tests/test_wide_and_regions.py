"""CPU checks of two host-side decompositions the library relies on (no GPU):

* max_xcell > 64: slots [64k, 64k + 64) of a CellArray are themselves a CellArray with the same cell stride, so a per-slot
  kernel may be launched once per chunk with shifted pointers.  Checked on the oracle: advecting /
  interpolating chunk views one at a time gives exactly what the full arrays give.
* jp_advect_region: the shell / interior brick partition (justpic/jl_b200/csrc/jp_advect_tile.cuh).  The predicate is restated
  here and checked for the property the halo overlap needs: every cell update_cell_halo! reads (layers 1, n-2) or rewrites
  (layers 0, n-1) lies in a shell brick, and the two regions partition the bricks."""
import itertools

import numpy as np
import pytest

from oracle.oracle import Oracle
from tests.problems import cfl_dt, make_grids, stream_velocity, vertex_field_linear


@pytest.mark.parametrize("ndim,n,S", [(2, (9, 7), 80), (3, (5, 4, 3), 150)])
def test_slot_chunks_are_cellarrays(ndim, n, S):
    gr = make_grids(n, ndim, True)
    full = Oracle(gr.xvi, gr.xci, gr.xi_vel, S, True)
    co, idx = full.init_particles(S - 10, 3)
    V = stream_velocity(gr); dt = cfl_dt(gr, V, 0.6)
    T = vertex_field_linear(gr) + 0.3
    ref = [c.copy() for c in co]; pT_ref = np.zeros_like(co[0])
    full.advect(ref, idx, 1, 0.5, V, dt)
    full.grid2particle(ref, idx, pT_ref, T)
    got = [c.copy() for c in co]; pT = np.zeros_like(co[0])
    for s0 in range(0, S, 64):
        s1 = min(S, s0 + 64)
        chunk = Oracle(gr.xvi, gr.xci, gr.xi_vel, s1 - s0, True)
        views = [c[s0:s1] for c in got]                       # contiguous: (slot, cell) at cell + slot * C
        assert all(v.flags.c_contiguous for v in views)
        chunk.advect(views, idx[s0:s1], 1, 0.5, V, dt)
        chunk.grid2particle(views, idx[s0:s1], pT[s0:s1], T)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(pT, pT_ref, equal_nan=True)


def _shell(b0, ext, n):
    """jp_advect_tile.cuh: brick [b0, min(b0 + ext, n)) per dimension holds a cell of layers {0, 1, n-2, n-1}."""
    return any(b0[d] <= 1 or min(b0[d] + ext[d], n[d]) >= n[d] - 1 for d in range(len(n)))


@pytest.mark.parametrize("n", [(70, 19), (33, 9), (256, 256), (40, 9, 7), (70, 11, 6), (72, 14, 10), (256, 256, 256), (257, 257, 257), (5, 4, 3)])
def test_region_partition_covers_the_exchanged_layers(n):
    ext = (32, 8) if len(n) == 2 else (32, 4, 2)             # AdvTile<N>::TX, TY, TZ
    nb = [-(-n[d] // ext[d]) for d in range(len(n))]
    n_shell = n_int = 0
    for b in itertools.product(*[range(k) for k in nb]):
        b0 = [b[d] * ext[d] for d in range(len(n))]
        sh = _shell(b0, ext, n)
        n_shell += sh; n_int += not sh
        if not sh:                                           # an interior brick holds no cell of the four outer layers
            for d in range(len(n)):
                lo, hi = b0[d], min(b0[d] + ext[d], n[d])
                assert lo >= 2 and hi <= n[d] - 2
    assert n_shell + n_int == int(np.prod(nb)) and n_shell > 0
    if n == (256, 256, 256):
        assert 0.27 < n_shell / (n_shell + n_int) < 0.30     # 28.5 % of the bricks travel first
