"""Synthetic problem set-ups shared by the tests (mirrors the reference's
scripts/tests: unit box, stream-function velocity, staggered grids with ghost
nodes; scripts/temperature_advection3D.jl:22-62, test/test_2D.jl:54-58)."""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

from justpic.jl_b200.api import LinRange, StepRange, expand_range


def make_grids(n, ndim, uniform=True, L=1.0, stretch=0.0, exact=False):
    """n cells per dim (int or tuple).  uniform -> LinRange grids (range path);
    else array grids (vector path), optionally stretched (non-uniform)."""
    ns = (n,) * ndim if isinstance(n, int) else tuple(n)
    xv, xc, xg = [], [], []
    for d in range(ndim):
        nd = ns[d]
        if uniform:
            R = StepRange if exact else LinRange      # exact: Julia range() (correctly rounded entries)
            v = R(0.0, L, nd + 1)
            dx = v[1] - v[0]
            c = R(0.0 + dx / 2, L - dx / 2, nd)
            g = expand_range(c)
        else:
            xi = np.asarray(LinRange(0.0, 1.0, nd + 1))
            v = L * (xi + stretch * np.sin(2 * math.pi * xi) / (2 * math.pi))
            c = 0.5 * (v[1:] + v[:-1])
            g = np.concatenate(([c[0] - (c[1] - c[0])], c, [c[-1] + (c[-1] - c[-2])]))
        xv.append(v); xc.append(c); xg.append(g)
    grid_vel = []
    for comp in range(ndim):
        grid_vel.append(tuple(xv[d] if d == comp else xg[d] for d in range(ndim)))
    asnp = lambda t: tuple(np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in t)
    xi_vel = tuple(asnp(g) for g in grid_vel)
    # centres are DERIVED from the velocity grids, as init_particles does
    # (src/Particles/particles_utils.jl:56-70): xci = interior of the ghosted vectors
    xci = (xi_vel[1][0][1:-1].copy(), xi_vel[0][1][1:-1].copy()) + ((xi_vel[0][2][1:-1].copy(),) if ndim == 3 else ())
    return SimpleNamespace(ndim=ndim, n=ns, uniform=uniform, grid_vel=tuple(grid_vel),
                           xvi=asnp(xv), xci=xci, xi_vel=xi_vel)


def stream_velocity(gr, amp=250.0):
    """vx = amp sin(pi x) cos(pi z|y), v_last = -amp cos(pi x) sin(pi z|y), (vy = 0 in 3D).
    Arrays are returned with shape ([nz,] ny, nx) = Julia (nx, ny[, nz]) column-major."""
    N = gr.ndim
    V = []
    for comp in range(N):
        gx = gr.xi_vel[comp]
        x = gx[0]
        z = gx[N - 1]
        if N == 2:
            X, Z = x[None, :], z[:, None]
            shape = (len(gx[1]), len(gx[0]))
        else:
            X, Z = x[None, None, :], z[:, None, None]
            shape = (len(gx[2]), len(gx[1]), len(gx[0]))
        if comp == 0:
            v = amp * np.sin(math.pi * X) * np.cos(math.pi * Z)
        elif comp == N - 1:
            v = -amp * np.cos(math.pi * X) * np.sin(math.pi * Z)
        else:
            v = np.zeros(shape)
        V.append(np.ascontiguousarray(np.broadcast_to(v, shape), dtype=np.float64))
    return V


def rotation_velocity(gr, w=math.pi * 1e-5):
    """solid rotation about the box centre (test/test_2D.jl:58, scripts/rotating_circle.jl:19); 2D."""
    assert gr.ndim == 2
    V = []
    for comp in range(2):
        gx = gr.xi_vel[comp]
        X, Y = gx[0][None, :], gx[1][:, None]
        shape = (len(gx[1]), len(gx[0]))
        v = -w * (Y - 0.5) + 0 * X if comp == 0 else w * (X - 0.5) + 0 * Y
        V.append(np.ascontiguousarray(np.broadcast_to(v, shape), dtype=np.float64))
    return V


def cfl_dt(gr, V, cfl):
    dts = []
    for d in range(gr.ndim):
        m = np.abs(V[d]).max()
        dx = np.diff(gr.xvi[d]).min()
        if m > 0:
            dts.append(dx / m)
    return cfl * min(dts)


def vertex_field_linear(gr, axis=-1):
    """T = coordinate along `axis` at the vertices, shape ([nz+1,] ny+1, nx+1)."""
    N = gr.ndim
    axis = axis % N
    shape = tuple(len(gr.xvi[d]) for d in reversed(range(N)))
    sl = [None] * N
    sl[N - 1 - axis] = slice(None)
    return np.ascontiguousarray(np.broadcast_to(gr.xvi[axis][tuple(sl)], shape), dtype=np.float64)


def centre_field_linear(gr, axis=-1):
    N = gr.ndim
    axis = axis % N
    shape = tuple(len(gr.xci[d]) for d in reversed(range(N)))
    sl = [None] * N
    sl[N - 1 - axis] = slice(None)
    return np.ascontiguousarray(np.broadcast_to(gr.xci[axis][tuple(sl)], shape), dtype=np.float64)
