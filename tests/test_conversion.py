"""Array(...) / CuArray(...) layout conversion (src/CellArrays/conversion.jl:19-69,
ext/JustPICCUDAExt.jl:51-187): the oracle restatement against a literal column-major transcription of
`permutedims(data, (3, 2, 1))` on the CPU, and the device kernel (jp_cellarray_permute through the public
API) against the oracle on the GPU -- bit-exact (a permutation; Float32 conversion is IEEE
round-to-nearest on both sides).  Shapes follow test/test_save_load.jl:154-159."""
import numpy as np
import pytest

from oracle import oracle as O


def _julia_permutedims_321(data_CS1):
    """data[C, S, 1] (Julia column-major) -> data[1, S, C]; both given / returned as flat memory images."""
    Cn, S = data_CS1.shape
    out = np.empty(Cn * S, dtype=data_CS1.dtype)
    for c in range(Cn):
        for s in range(S):
            out[s + c * S] = data_CS1[c, s]            # dst[1, s, c] at (1-1) + 1*(s + S*c)
    return out


@pytest.mark.parametrize("shape", [(5, 3, 4), (7, 2, 3, 5), (1, 6, 2)])
def test_oracle_layout_is_permutedims_321(shape):
    rng = np.random.default_rng(0)
    a = rng.random(shape)                               # (S, [nz,] ny, nx): memory offset c + s*C
    S, Cn = shape[0], int(np.prod(shape[1:]))
    flat_dev = a.reshape(-1)
    data_CS1 = np.empty((Cn, S))
    for c in range(Cn):
        for s in range(S):
            data_CS1[c, s] = flat_dev[c + s * Cn]       # Julia data[c+1, s+1, 1]
    want = _julia_permutedims_321(data_CS1)
    got = O.cellarray_to_host_layout(a)
    assert got.shape == (*shape[1:], S)
    assert np.array_equal(got.reshape(-1), want)
    assert np.array_equal(O.cellarray_to_device_layout(got), a)
    assert O.cellarray_to_host_layout(a, np.float32).dtype == np.float32
    m = rng.random(shape) < 0.5
    assert O.cellarray_to_host_layout(m.astype(np.uint8)).dtype == np.bool_
    assert np.array_equal(O.cellarray_to_device_layout(O.cellarray_to_host_layout(m.astype(np.uint8))), m.astype(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(48, 9, 7, 11), (5, 33, 31), (64, 4, 4, 4), (33, 1, 70), (2, 17, 17, 18), (24, 40, 64)])
def test_cellarray_roundtrip_and_types(shape):
    import torch
    import justpic.jl_b200 as J
    rng = np.random.default_rng(1)
    a = rng.standard_normal(shape)
    a[rng.random(shape) < 0.2] = np.nan
    t = torch.from_numpy(a).cuda()
    h = J.Array(t)
    assert h.shape == (*shape[1:], shape[0]) and h.dtype == np.float64
    assert np.array_equal(h, O.cellarray_to_host_layout(a), equal_nan=True)
    h32 = J.Array(t, np.float32)
    assert h32.dtype == np.float32 and np.array_equal(h32, O.cellarray_to_host_layout(a, np.float32), equal_nan=True)
    back = J.CuArray(h)
    assert back.dtype == torch.float64 and tuple(back.shape) == shape
    assert np.array_equal(back.cpu().numpy(), a, equal_nan=True)
    b32 = J.CuArray(h, np.float32)
    assert b32.dtype == torch.float32 and np.array_equal(b32.cpu().numpy(), a.astype(np.float32), equal_nan=True)
    up = J.CuArray(h32, np.float64)                      # Float32 checkpoint -> Float64 device array
    assert up.dtype == torch.float64 and np.array_equal(up.cpu().numpy(), a.astype(np.float32).astype(np.float64), equal_nan=True)
    m = (rng.random(shape) < 0.5).astype(np.uint8)
    hm = J.Array(torch.from_numpy(m).cuda(), np.float32)  # index stays Bool whatever T is (conversion.jl:50-51)
    assert hm.dtype == np.bool_ and np.array_equal(hm, O.cellarray_to_host_layout(m))
    assert np.array_equal(J.CuArray(hm).cpu().numpy(), m)


@pytest.mark.gpu
@pytest.mark.parametrize("ndim", [2, 3])
def test_particles_and_phase_ratios_checkpoint_roundtrip(ndim):
    """test/test_save_load.jl:133-173 in spirit: Array(particles) -> CuArray(...) gives a container that
    continues the run exactly like the original."""
    import torch
    import justpic.jl_b200 as J
    from tests.problems import cfl_dt, make_grids, stream_velocity
    gr = make_grids(12 if ndim == 2 else (9, 7, 8), ndim, True)
    p = J.init_particles(J.CUDABackend, 12, 24, 6, *gr.grid_vel, seed=3)
    V = stream_velocity(gr); Vd = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in V]
    dt = cfl_dt(gr, V, 0.8)
    ph, = J.init_cell_arrays(p, 1)
    ph.copy_(torch.where(p.index > 0, 1.0 + (p.coords[0] < p.coords[-1]).double(), torch.zeros_like(ph)))
    J.advection(p, J.RungeKutta2(), Vd, dt); J.move_particles(p, (ph,)); J.inject_particles(p, (ph,))
    pr = J.PhaseRatios(J.CUDABackend, 2, gr.n)
    J.update_phase_ratios(pr, p, ph)
    hp, hph, hpr = J.Array(p), J.Array(ph), J.Array(pr)
    assert isinstance(hp, J.HostParticles) and hp.index.dtype == np.bool_
    for d in range(ndim):
        assert hp.coords[d].shape == (*reversed(gr.n), p.max_xcell)
        assert np.array_equal(hp.coords[d], O.cellarray_to_host_layout(p.coords[d].cpu().numpy()), equal_nan=True)
    assert np.array_equal(hp.index, O.cellarray_to_host_layout(p.index.cpu().numpy()))
    assert hpr.vertex.shape == (*(n + 1 for n in reversed(gr.n)), 2)
    assert np.array_equal(hpr.center, O.cellarray_to_host_layout(pr.center.cpu().numpy()))
    p2, ph2, pr2 = J.CuArray(hp), J.CuArray(hph), J.CuArray(hpr)
    assert isinstance(p2, J.Particles) and p2._ctx != p._ctx
    for f in ("center", "vertex", "Vx", "Vy", "Vz", "yz", "xz", "xy"):
        assert torch.equal(getattr(pr2, f), getattr(pr, f))
    for it in range(3):                                  # both containers continue identically
        for q, f in ((p, ph), (p2, ph2)):
            J.advection(q, J.RungeKutta2(), Vd, dt); J.move_particles(q, (f,)); J.inject_particles(q, (f,))
        for d in range(ndim):
            assert np.array_equal(p.coords[d].cpu().numpy(), p2.coords[d].cpu().numpy(), equal_nan=True)
        assert torch.equal(p.index, p2.index) and np.array_equal(ph.cpu().numpy(), ph2.cpu().numpy(), equal_nan=True)
    with pytest.raises(NotImplementedError):
        J.CuArray(hp, np.float32)
    with pytest.raises(TypeError):
        J.Array("not a cell array")
