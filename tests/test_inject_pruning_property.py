"""inject_particles!: the donor search of k_inject_sweep skips a neighbour cell when every particle it holds lies in the cell's
closed box and the box is farther from the new particle than the best donor so far (justpic_sm100a.cu, "inbox" pruning).  That is
only exact if, IN FLOATING POINT, the computed distance to the nearest point of the box never exceeds the computed distance to any
point inside the box -- every operation of distance() (src/Interpolations/utils.jl:9-19: subtract, square, add, sqrt) is monotonic
under round-to-nearest, which this test checks on adversarial inputs (points on faces / corners, one ulp inside, huge and tiny
scales), together with the key order that replaces the reference's serial "strictly smaller" scan."""
import numpy as np
import pytest


def dist(a, b):
    s = (a[0] - b[0]) * (a[0] - b[0])
    for d in range(1, len(a)):
        s = s + (a[d] - b[d]) * (a[d] - b[d])
    return np.sqrt(s)


@pytest.mark.parametrize("N", [2, 3])
@pytest.mark.parametrize("scale", [1.0, 1e-9, 3e7])
def test_box_distance_is_a_lower_bound_in_floating_point(N, scale):
    rng = np.random.default_rng(17 + N)
    worst = 0
    for it in range(1200):
        lo = np.float64(scale) * rng.uniform(-1, 1, N)
        hi = lo + np.float64(scale) * rng.uniform(1e-3, 1, N) * rng.choice([1.0, 1e-6], N)
        pn = lo + (hi - lo) * rng.uniform(-2, 3, N)                       # inside, beside, diagonal to the box
        if it % 5 == 0:                                                    # exactly on a face / corner plane of the box
            k = rng.integers(0, N); pn[k] = rng.choice([lo[k], hi[k]])
        bx = np.where(pn < lo, lo, np.where(pn > hi, hi, pn))              # nearest point of the box (the kernel's clamp)
        dbox = dist(bx, pn)
        for _ in range(12):
            q = lo + (hi - lo) * rng.uniform(0, 1, N)
            m = rng.integers(0, 4, N)                                      # push coordinates onto faces / one ulp inside
            q = np.where(m == 0, lo, np.where(m == 1, hi, np.where(m == 2, np.nextafter(lo, hi), q)))
            q = np.minimum(np.maximum(q, lo), hi)
            assert dbox <= dist(q, pn), (lo, hi, pn, q)
            worst += 1
    assert worst > 12000


def test_key_order_equals_serial_strictly_smaller_scan():
    """min over (distance, visiting order) == the candidate kept by a serial scan that replaces only on a strictly smaller
    distance (index_min_distance, src/Particles/injection.jl:330-393), including exact ties."""
    rng = np.random.default_rng(3)
    for it in range(2000):
        n = int(rng.integers(1, 40))
        d = rng.choice([0.25, 0.5, 0.75, 1.0, rng.uniform()], n)          # many exact ties
        best, keep = np.inf, -1
        for i in range(n):                                                 # the reference's scan, in visiting order
            if d[i] < best:
                best, keep = d[i], i
        order = rng.permutation(n)                                         # the kernel evaluates candidates in any order
        kd, ko = np.inf, 1 << 30
        for i in order:
            if d[i] < kd or (d[i] == kd and i < ko):
                kd, ko = d[i], i
        assert ko == keep
