// jp_core.h -- per-particle / per-cell arithmetic of the JustPIC hot path.
//
// Everything here is `__host__ __device__` so the CUDA kernels
// (justpic_sm100a.cu) and the CPU emulation used by the no-GPU tests
// (tests/emul/jp_emul.cpp) compile the *same* source.  Floating-point contract:
// nvcc -fmad=false / g++ -ffp-contract=off, fma() explicit exactly where the
// reference has muladd/fma/@muladd, IEEE division and sqrt.  Reference
// file:line citations are relative to /root/reference.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define JP_HD __host__ __device__ __forceinline__
#else
#define JP_HD inline
#endif

#define JP_MAX_ARGS 16
#define JP_MAX_SLOTS 64          // occupancy-word kernels (one 64-bit mask per cell)
#define JP_MAX_SLOTS_WIDE 1024   // max_xcell > 64: slot-chunked launches + literal per-cell move / inject kernels
#define JP_MAX_PHASES 32

// Grid tables.  Pointers are device pointers in the library and host pointers
// in the emulation.  Derived fields are filled by jp_grid_derive() on the host.
struct JpGrid {
    int32_t ndim, n[3], S, uniform;
    int64_t C;
    const double *xv[3];       // vertices, n+1
    const double *xc[3];       // centres, n
    const double *xvel[3][3];  // xvel[comp][dim]
    int32_t nvel[3][3];
    // ---- derived ----
    // kind of xvel[comp][dim]: 1 = bitwise identical to xv[dim] ("V"), 2 = bitwise
    // identical to the ghosted-centre vector xg[dim] ("G"), 0 = anything else.
    int32_t vkind[3][3];
    const double *xg[3];       // canonical ghosted-centre vector per dim (n+2) or null
    const double *ixv[3];      // 1/diff(xv)   (n)      -- non-uniform fast path
    const double *ixg[3];      // 1/diff(xg)   (n+1)
    double inv_dv[3], inv_dg[3];  // uniform: 1/(x[1]-x[0])
    int32_t fast;              // 1: every (comp,dim) is V or G and xg[i] < xv[i] < xg[i+1] for all i
    // 1: checked on the host for EVERY entry, bit for bit:  xv[d][i] == fma(i, aff_dv[d], aff_v0[d])  and
    //    xg[d][j] == fma(j, aff_dg[d], aff_g0[d])  -- the tiled advection kernel then regenerates grid
    //    coordinates in registers instead of loading them (same bits by construction)
    int32_t affine;
    double aff_v0[3], aff_dv[3], aff_g0[3], aff_dg[3];
    double dom_lo[3], dom_hi[3];   // xv[d][0], xv[d][n]  (kernel-parameter constants instead of per-thread loads)
    double dxv0[3];                // xv[d][1] - xv[d][0]  (the scalar spacing of range grids)
    double dxg0[3];                // xg[d][1] - xg[d][0]  (the same for the ghosted-centre vectors)
    double inv_dmin_v[3];          // inv(grid_size(xvi)) = 1 / abs(minimum(diff(xv)))  (grid2particle_flip!)
    // 1: range grid whose vertices were checked on the host to be affine within 1e-6 dx (and dx >> ulp of the
    //    coordinates): jp_classify_fast may decide particles that are clearly inside a cell (see there)
    int32_t cls_fast;
    // advection! split for the halo overlap (jp_advect_region): 0 = every cell, 1 = only bricks that hold a cell of the two
    // outermost cell layers (what update_cell_halo! reads and rewrites), 2 = the other bricks
    int32_t region;
    // 1: every live particle lies strictly inside its storage cell (the state move_particles! / init / inject / clean leave);
    // the tiled advection kernel then skips the re-centring of its first interpolation.  Set per launch from the context's state.
    int32_t bucketed;
};

struct JpArgs {               // particle fields carried along by move/inject/clean
    double *a[JP_MAX_ARGS];
    int32_t n;
};

// ---------------------------------------------------------------------------
// spacing accessor: scalar x[1]-x[0] for range grids
// (src/Particles/particles_utils.jl:137-140), diff(x)[i] for array grids (:76-79)
JP_HD double jp_d_of(const double *x, int uniform, int i0) {
    return uniform ? x[1] - x[0] : x[i0 + 1] - x[i0];
}

// find_parent_cell_bisection (src/Utils.jl:117-130); 1-based arithmetic kept literally.
JP_HD int jp_bisect1(double px, const double *x, int len, int seed1) {
    int lo = 1, hi = len, seed = seed1;
    for (;;) {
        if (x[seed - 1] <= px && px <= x[seed]) return seed;
        if (x[seed - 1] < px) { lo = seed; seed = (hi + seed) / 2; }
        else                  { hi = seed; seed = (lo + seed) / 2; }
    }
}

// lerp (src/Interpolations/ndlerp.jl:11-15)
JP_HD double jp_lerp1(double t, double v0, double v1) { return fma(t, v1, fma(-t, v0, v0)); }
template <int N> JP_HD double jp_lerp(const double *v, const double *t);
template <> JP_HD double jp_lerp<2>(const double *v, const double *t) {
    return jp_lerp1(t[1], jp_lerp1(t[0], v[0], v[1]), jp_lerp1(t[0], v[2], v[3]));
}
template <> JP_HD double jp_lerp<3>(const double *v, const double *t) {
    double a = jp_lerp1(t[1], jp_lerp1(t[0], v[0], v[1]), jp_lerp1(t[0], v[2], v[3]));
    double b = jp_lerp1(t[1], jp_lerp1(t[0], v[4], v[5]), jp_lerp1(t[0], v[6], v[7]));
    return jp_lerp1(t[2], a, b);
}

// corners of an N-d column-major array at linear base b
// (extract_field_corners src/Advection/advection.jl:3-27, field_corners
// src/Interpolations/utils.jl:98-118)
template <int N> JP_HD void jp_corners(const double *F, int64_t b, int64_t s1, int64_t s2, double *v) {
    v[0] = F[b]; v[1] = F[b + 1]; v[2] = F[b + s1]; v[3] = F[b + s1 + 1];
    if (N == 3) { v[4] = F[b + s2]; v[5] = F[b + s2 + 1]; v[6] = F[b + s2 + s1]; v[7] = F[b + s2 + s1 + 1]; }
}

// ---------------------------------------------------------------------------
// Velocity interpolation, literal restatement of interp_velocity2particle
// (src/Particles/Advection/advection.jl:93-148) incl. check_local_limits
// (src/Advection/advection.jl:39-47) and normalize_coordinates
// (src/Interpolations/utils.jl:54-58).  cell1 = 1-based storage cell (seed).
template <int N>
JP_HD void jp_interp_velocity_literal(const JpGrid &g, const double *const *V, const double *p, const int *cell1, double *vout) {
    for (int c = 0; c < N; c++) {
        bool ok = true;
        for (int d = 0; d < N; d++) {
            const double *x = g.xvel[c][d];
            if (!(x[0] <= p[d] && p[d] <= x[g.nvel[c][d] - 1])) { ok = false; break; }
        }
        if (!ok) { vout[c] = INFINITY; continue; }
        int idx[3] = {1, 1, 1};
        double t[3];
        for (int d = 0; d < N; d++) {
            const double *x = g.xvel[c][d];
            idx[d] = jp_bisect1(p[d], x, g.nvel[c][d], cell1[d]);
            double dx = jp_d_of(x, g.uniform, idx[d] - 1);
            t[d] = (p[d] - x[idx[d] - 1]) * (1.0 / dx);
        }
        const int64_t s1 = g.nvel[c][0], s2 = (int64_t)g.nvel[c][0] * g.nvel[c][1];
        const int64_t b = (idx[0] - 1) + s1 * (idx[1] - 1) + (N == 3 ? s2 * (idx[2] - 1) : 0);
        double v[8];
        jp_corners<N>(V[c], b, s1, s2, v);
        vout[c] = jp_lerp<N>(v, t);
    }
}

// ---------------------------------------------------------------------------
// advection_MQS! / advection_LinP! interpolants (src/Interpolations/MQS.jl,
// src/Particles/Advection/advection_LinP.jl:96-391, advection_MQS.jl:96-124): the linear
// interpolant of interp_velocity2particle plus a correction when the interpolation cell is interior
// (1 < idx < size(F) - 1 in every direction).  Literal, quirks included (listed in DESIGN.md).
// The stencil values are read through an accessor A(i1, j1, k1) (1-based GLOBAL node indices of the velocity array): the global array
// itself (JpGlobalAcc) or the shared-memory tile of the tiled advection kernel (jp_advect_tile.cuh) -- same arithmetic either way.
template <int N> struct JpGlobalAcc {
    const double *F; const int32_t *nF;
    JP_HD double operator()(int i1, int j1, int k1) const {
        return F[(i1 - 1) + (int64_t)nF[0] * ((j1 - 1) + (N == 3 ? (int64_t)nF[1] * (k1 - 1) : 0))];
    }
};
template <int N, class Acc> JP_HD double jp_Fat(const Acc &F, const int32_t *nF, int i1, int j1, int k1) { (void)nF; return F(i1, j1, k1); }
// quadratic correction of one edge (v0e, v1e) at tq with the outer nodes on either side
JP_HD double jp_mqs_edge(double v0e, double v1e, double tq, double outer_lo, double outer_hi) {
    const double l = jp_lerp1(tq, v0e, v1e);
    const bool low = tq < 0.5;
    const double a = low ? outer_lo : v0e, b = low ? v0e : v1e, c = low ? v1e : outer_hi;
    return l + (0.5 * ((tq - 0.5) * (tq - 0.5))) * (fma(-2.0, b, a) + c);
}
template <int N, class Acc> JP_HD double jp_mqs(const Acc &F, const int32_t *nF, int comp, const int *idx1, const double *v, const double *t) {
    const int i = idx1[0], j = idx1[1], k = N == 3 ? idx1[2] : 1;
    if (N == 2) {
        if (comp == 0)
            return jp_lerp1(t[1], jp_mqs_edge(v[0], v[1], t[0], jp_Fat<N>(F, nF, i - 1, j, 1), jp_Fat<N>(F, nF, i + 2, j, 1)),
                            jp_mqs_edge(v[2], v[3], t[0], jp_Fat<N>(F, nF, i - 1, j + 1, 1), jp_Fat<N>(F, nF, i + 2, j + 1, 1)));
        return jp_lerp1(t[0], jp_mqs_edge(v[0], v[2], t[1], jp_Fat<N>(F, nF, i, j - 1, 1), jp_Fat<N>(F, nF, i, j + 2, 1)),
                        jp_mqs_edge(v[1], v[3], t[1], jp_Fat<N>(F, nF, i + 1, j - 1, 1), jp_Fat<N>(F, nF, i + 1, j + 2, 1)));
    }
    double f[2];
    for (int h = 0; h < 2; h++) {
        if (comp == 0) {            // MQS-x of v[1:4] / v[5:8], outer nodes of BOTH taken in plane k (MQS.jl:62-70, :80-108)
            const double *w = v + 4 * h;
            f[h] = jp_lerp1(t[1], jp_mqs_edge(w[0], w[1], t[0], jp_Fat<N>(F, nF, i - 1, j, k), jp_Fat<N>(F, nF, i + 2, j, k)),
                            jp_mqs_edge(w[2], w[3], t[0], jp_Fat<N>(F, nF, i - 1, j + 1, k), jp_Fat<N>(F, nF, i + 2, j + 1, k)));
        } else if (comp == 1) {     // MQS-y (MQS.jl:110-133)
            const double *w = v + 4 * h;
            f[h] = jp_lerp1(t[0], jp_mqs_edge(w[0], w[2], t[1], jp_Fat<N>(F, nF, i, j - 1, k), jp_Fat<N>(F, nF, i, j + 2, k)),
                            jp_mqs_edge(w[1], w[3], t[1], jp_Fat<N>(F, nF, i + 1, j - 1, k), jp_Fat<N>(F, nF, i + 1, j + 2, k)));
        } else {                    // MQS-z: front (v1,v2,v5,v6) / back (v3,v4,v7,v8) with (t1,t3), same j, correction along x (MQS.jl:73-78, :135-158)
            const double w[4] = {v[2 * h], v[2 * h + 1], v[4 + 2 * h], v[4 + 2 * h + 1]};
            f[h] = jp_lerp1(t[2], jp_mqs_edge(w[0], w[1], t[0], jp_Fat<N>(F, nF, i - 1, j, k), jp_Fat<N>(F, nF, i + 2, j, k)),
                            jp_mqs_edge(w[2], w[3], t[0], jp_Fat<N>(F, nF, i - 1, j, k + 1), jp_Fat<N>(F, nF, i + 2, j, k + 1)));
        }
    }
    return comp == 2 ? jp_lerp1(t[1], f[0], f[1]) : jp_lerp1(t[2], f[0], f[1]);
}
JP_HD int jp_clampi(int x, int lo, int hi) { return x > hi ? hi : (x < lo ? lo : x); }
template <int N, class Acc> JP_HD double jp_linp(const Acc &F, const int32_t *nF, int comp, const int *idx1, const double *xc, const double *dxi,
                                               const double *p, double VL) {
    int ijk[3] = {idx1[0], idx1[1], N == 3 ? idx1[2] : 1};
    ijk[comp] += p[comp] > xc[comp] + dxi[comp] / 2 ? 1 : 0;                  // offset_LinP (advection_LinP.jl:345-347)
    // augment_offset (:364-391): three offsets (-1, 0, 1) along `comp`; the two transverse directions take
    // (0|1): rows (1,1,1), (2,2,2), (3,1,3), (4,2,4) of (offset_i, offset_j, offset_k)
    const int ta = comp == 0 ? 1 : 0, tb = comp == 2 ? 1 : 2;                // transverse dims, lower first
    const int nrow = N == 2 ? 2 : 4;
    double av[8];
    for (int r = 0; r < nrow; r++) {
        // bit pattern of the tables: offset_j rows (0,1,0,1) [used as rows 1,2,1,2], offset_k / offset_i rows (0,0,1,1)
        int o[3] = {0, 0, 0};
        if (comp == 0) { o[1] = r & 1; o[2] = r >> 1; }
        else if (comp == 1) { o[0] = r & 1; o[2] = r >> 1; }
        else { o[0] = r >> 1; o[1] = r & 1; }
        (void)ta; (void)tb;
        double f[3];
        for (int m = 0; m < 3; m++) {
            int q[3] = {ijk[0] + o[0], ijk[1] + o[1], ijk[2] + o[2]};
            q[comp] = ijk[comp] + (m - 1);
            f[m] = jp_Fat<N>(F, nF, jp_clampi(q[0], 1, nF[0]), jp_clampi(q[1], 1, nF[1]), N == 3 ? jp_clampi(q[2], 1, nF[2]) : 1);
        }
        av[2 * r] = (f[0] + f[1]) / 2;
        av[2 * r + 1] = (f[2] + f[1]) / 2;
    }
    double FP[8];
    if (comp == 0) { for (int q = 0; q < 2 * nrow; q++) FP[q] = av[q]; }
    else {                                                                  // swap_F (:212-216, :335-340)
        FP[0] = av[0]; FP[1] = av[2]; FP[2] = av[1]; FP[3] = av[3];
        if (N == 3) { FP[4] = av[4]; FP[5] = av[6]; FP[6] = av[5]; FP[7] = av[7]; }
    }
    double tP[3];
    for (int d = 0; d < N; d++) {
        double x = xc[d];
        if (d == comp) x = xc[d] + ((double)(1 - 2 * (p[d] < xc[d] + dxi[d] / 2 ? 1 : 0)) * dxi[d]) / 2;   // correct_xci_to_pressure_point
        tP[d] = (p[d] - x) * (1.0 / dxi[d]);
    }
    const double VP = jp_lerp<N>(FP, tP);
    const double A = 2.0 / 3.0;
    return A * VL + (1 - A) * VP;
}

// interp_velocity2particle_LinP / _MQS: INTERP = 1 LinP, 2 MQS (0 = the linear interpolant)
template <int N, int INTERP>
JP_HD void jp_interp_velocity_hi(const JpGrid &g, const double *const *V, const double *p, const int *cell1, double *vout) {
    for (int c = 0; c < N; c++) {
        bool ok = true;
        for (int d = 0; d < N; d++) {
            const double *x = g.xvel[c][d];
            if (!(x[0] <= p[d] && p[d] <= x[g.nvel[c][d] - 1])) { ok = false; break; }
        }
        if (!ok) { vout[c] = INFINITY; continue; }
        int idx[3] = {1, 1, 1};
        double t[3], xcn[3], dxi[3];
        bool interior = true;
        for (int d = 0; d < N; d++) {
            const double *x = g.xvel[c][d];
            idx[d] = jp_bisect1(p[d], x, g.nvel[c][d], cell1[d]);
            xcn[d] = x[idx[d] - 1];
            dxi[d] = jp_d_of(x, g.uniform, idx[d] - 1);
            t[d] = (p[d] - xcn[d]) * (1.0 / dxi[d]);
            interior = interior && 1 < idx[d] && idx[d] < g.nvel[c][d] - 1;
        }
        const int64_t s1 = g.nvel[c][0], s2 = (int64_t)g.nvel[c][0] * g.nvel[c][1];
        double v[8];
        jp_corners<N>(V[c], (idx[0] - 1) + s1 * (idx[1] - 1) + (N == 3 ? s2 * (idx[2] - 1) : 0), s1, s2, v);
        const double VL = jp_lerp<N>(v, t);
        if (INTERP == 0 || !interior) vout[c] = VL;
        else {
            const JpGlobalAcc<N> acc = {V[c], g.nvel[c]};
            if (INTERP == 2) vout[c] = jp_mqs<N>(acc, g.nvel[c], c, idx, v, t);
            else vout[c] = jp_linp<N>(acc, g.nvel[c], c, idx, xcn, dxi, p, VL);
        }
    }
}

// Fast path.  Preconditions (g.fast): every xvel[comp][dim] is the vertex
// vector ("V") or the ghosted-centre vector ("G") of that dim, and
// xg[i] < xv[i] < xg[i+1] for every vertex i.  For a particle STRICTLY inside a
// vertex cell found within seed-1..seed+1 the bisection result is unique, so it
// can be produced by comparisons alone; the index of a G vector follows from one
// more comparison.  Any tie, NaN, >1-cell displacement or out-of-domain
// coordinate returns false and the caller runs the literal routine, so results
// are bitwise identical to the literal one in every case.
template <int N, bool UNIFORM>
JP_HD bool jp_interp_velocity_fast(const JpGrid &g, const double *const *V, const double *p, const int *cell1, double *vout) {
    int iv[3], ig[3];
    double tv[3], tg[3];
#pragma unroll
    for (int d = 0; d < N; d++) {
        const double *xv = g.xv[d];
        const double pd = p[d];
        int i = cell1[d] - 1;
        double a = xv[i], b = xv[i + 1];
        if (!(a < pd && pd < b)) {
            if (pd > b) { i += 1; if (i >= g.n[d]) return false; }
            else if (pd < a) { i -= 1; if (i < 0) return false; }
            else return false;                       // tie or NaN
            a = xv[i]; b = xv[i + 1];
            if (!(a < pd && pd < b)) return false;
        }
        iv[d] = i;
        tv[d] = (pd - a) * (UNIFORM ? g.inv_dv[d] : g.ixv[d][i]);
        const double *xg = g.xg[d];
        if (xg) {
            const double m = xg[i + 1];
            if (pd == m) return false;               // tie on a G node
            const int j = pd < m ? i : i + 1;
            ig[d] = j;
            tg[d] = (pd - (pd < m ? xg[i] : m)) * (UNIFORM ? g.inv_dg[d] : g.ixg[d][j]);
        } else { ig[d] = i; tg[d] = tv[d]; }
    }
#pragma unroll
    for (int c = 0; c < N; c++) {
        int idx[3] = {0, 0, 0};
        double t[3];
#pragma unroll
        for (int d = 0; d < N; d++) {
            const bool isv = g.vkind[c][d] == 1;
            idx[d] = isv ? iv[d] : ig[d];
            t[d] = isv ? tv[d] : tg[d];
        }
        const int64_t s1 = g.nvel[c][0], s2 = (int64_t)g.nvel[c][0] * g.nvel[c][1];
        const int64_t b = idx[0] + s1 * idx[1] + (N == 3 ? s2 * idx[2] : 0);
        double v[8];
        jp_corners<N>(V[c], b, s1, s2, v);
        vout[c] = jp_lerp<N>(v, t);
    }
    return true;
}

template <int N, bool FAST, bool UNIFORM>
JP_HD void jp_interp_velocity(const JpGrid &g, const double *const *V, const double *p, const int *cell1, double *vout) {
    if (FAST) { if (jp_interp_velocity_fast<N, UNIFORM>(g, V, p, cell1, vout)) return; }
    jp_interp_velocity_literal<N>(g, V, p, cell1, vout);
}

// Integrator stages (src/Advection/Euler.jl:1-6, RK2.jl:1-38):
// @muladd a + s*c*dt*v -> muladd((s*c)*dt, v, a)   (MuladdMacro 0.2.4)
// RK4 (src/Particles/Advection/RK4.jl:1-19): unfused, left to right.
template <int N, int SCHEME, bool FAST, bool UNIFORM>
JP_HD void jp_advect_particle(const JpGrid &g, double alpha, const double *const *V, double dt,
                              const int *cell1, const double *p0, double *pout) {
    double k1[3], k2[3], q[3];
    jp_interp_velocity<N, FAST, UNIFORM>(g, V, p0, cell1, k1);
    if (SCHEME == 0) {
        const double c = 1.0 * dt;
        for (int d = 0; d < N; d++) pout[d] = fma(c, k1[d], p0[d]);
    } else if (SCHEME == 1) {
        const double c = (1.0 * alpha) * dt;
        for (int d = 0; d < N; d++) q[d] = fma(c, k1[d], p0[d]);
        jp_interp_velocity<N, FAST, UNIFORM>(g, V, q, cell1, k2);
        if (alpha == 0.5) {
            for (int d = 0; d < N; d++) pout[d] = fma(1.0 * dt, k2[d], p0[d]);
        } else {
            const double b = 0.5 * (1.0 / alpha), a = 1.0 - b;
            for (int d = 0; d < N; d++) pout[d] = fma(1.0 * dt, fma(b, k2[d], a * k1[d]), p0[d]);
        }
    } else {
        double k3[3], k4[3];
        for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k1[d] / 2;
        jp_interp_velocity<N, FAST, UNIFORM>(g, V, q, cell1, k2);
        for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k2[d] / 2;
        jp_interp_velocity<N, FAST, UNIFORM>(g, V, q, cell1, k3);
        for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k3[d];
        jp_interp_velocity<N, FAST, UNIFORM>(g, V, q, cell1, k4);
        for (int d = 0; d < N; d++) pout[d] = p0[d] + dt * (((k1[d] + 2 * k2[d]) + 2 * k3[d]) + k4[d]) / 6;
    }
}

// advection_LinP! / advection_MQS!: same integrators (advect_particle with interpolation_fn,
// src/Particles/Advection/Euler.jl, RK2.jl:28-54, RK4.jl), other interpolant
template <int N, int SCHEME, int INTERP>
JP_HD void jp_advect_particle_hi(const JpGrid &g, double alpha, const double *const *V, double dt,
                                 const int *cell1, const double *p0, double *pout) {
    double k1[3], k2[3], q[3];
    jp_interp_velocity_hi<N, INTERP>(g, V, p0, cell1, k1);
    if (SCHEME == 0) {
        const double c = 1.0 * dt;
        for (int d = 0; d < N; d++) pout[d] = fma(c, k1[d], p0[d]);
    } else if (SCHEME == 1) {
        const double c = (1.0 * alpha) * dt;
        for (int d = 0; d < N; d++) q[d] = fma(c, k1[d], p0[d]);
        jp_interp_velocity_hi<N, INTERP>(g, V, q, cell1, k2);
        if (alpha == 0.5) {
            for (int d = 0; d < N; d++) pout[d] = fma(1.0 * dt, k2[d], p0[d]);
        } else {
            const double b = 0.5 * (1.0 / alpha), a = 1.0 - b;
            for (int d = 0; d < N; d++) pout[d] = fma(1.0 * dt, fma(b, k2[d], a * k1[d]), p0[d]);
        }
    } else {
        double k3[3], k4[3];
        for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k1[d] / 2;
        jp_interp_velocity_hi<N, INTERP>(g, V, q, cell1, k2);
        for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k2[d] / 2;
        jp_interp_velocity_hi<N, INTERP>(g, V, q, cell1, k3);
        for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k3[d];
        jp_interp_velocity_hi<N, INTERP>(g, V, q, cell1, k4);
        for (int d = 0; d < N; d++) pout[d] = p0[d] + dt * (((k1[d] + 2 * k2[d]) + 2 * k3[d]) + k4[d]) / 6;
    }
}

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11).
JP_HD void jp_philox4x32_10(uint32_t *ctr, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1, n3 = (uint32_t)p0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
JP_HD double jp_u01(uint32_t hi, uint32_t lo) {
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * 0x1.0p-53;
}
// three uniforms in [0,1) for (seed, purpose, step, cell, slot)
JP_HD void jp_rand3(uint64_t seed, uint32_t purpose, uint32_t step, uint32_t cell, uint32_t slot, double *r) {
    uint32_t a[4] = {cell, slot, purpose, (step << 1) | 0u};
    uint32_t b[4] = {cell, slot, purpose, (step << 1) | 1u};
    jp_philox4x32_10(a, (uint32_t)seed, (uint32_t)(seed >> 32));
    jp_philox4x32_10(b, (uint32_t)seed, (uint32_t)(seed >> 32));
    r[0] = jp_u01(a[0], a[1]); r[1] = jp_u01(a[2], a[3]); r[2] = jp_u01(b[0], b[1]);
}

// ---------------------------------------------------------------------------
// small helpers
template <int N> JP_HD void jp_cell_ijk(const JpGrid &g, int64_t c, int *ci) {
    ci[0] = (int)(c % g.n[0]);
    if (N == 2) { ci[1] = (int)(c / g.n[0]); ci[2] = 0; }
    else { ci[1] = (int)((c / g.n[0]) % g.n[1]); ci[2] = (int)(c / ((int64_t)g.n[0] * g.n[1])); }
}
template <int N> JP_HD int64_t jp_cell_lin(const JpGrid &g, const int *ci) {
    return ci[0] + (int64_t)g.n[0] * (ci[1] + (N == 3 ? (int64_t)g.n[1] * ci[2] : 0));
}

// ---- move_particles! classification of ONE live particle (shared by k_move_classify3 and the
// advection -> move hand-off of the tiled advection kernel, so both produce the same byte from the same
// comparisons).  am/a/b/bp: the four vertices around the storage cell per dimension (NaN outside the
// grid, so every comparison with them fails).  Returns
//   JP_CLS_STAY         strictly inside its cell (isincell, upper edge fl(a + dx), move_safe.jl:93)
//   0..26               left for the neighbour with that direction code (x fastest); 2-D codes use dz = 0
//   JP_CODE_DELETE      left the domain (indomain, move_safe.jl:96-103)
//   JP_CLS_CPLX + r     the planner cannot express it (r = 1: on a vertex / more than one cell away,
//                       2: bisects back into its own cell (ulp gap), 3: fails isincell in its destination)
#define JP_CODE_DELETE 27
#define JP_CLS_STAY 28
#define JP_CLS_CPLX 28
template <int N>
JP_HD int jp_classify_particle(const JpGrid &g, const double *am, const double *a, const double *b, const double *bp, const double *p) {
    bool in = true, indom = true, near = true, dest_ok = true;
    int dv[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < N; d++) {
        const double pd = p[d];
        const double dx0 = g.uniform ? g.dxv0[d] : b[d] - a[d];
        in = in & (a[d] < pd) & (pd < a[d] + dx0);
        indom = indom & (g.dom_lo[d] < pd) & (pd < g.dom_hi[d]);
        double lower, dxd;
        if (a[d] < pd && pd < b[d]) { dv[d] = 0; lower = a[d]; dxd = dx0; }
        else if (am[d] < pd && pd < a[d]) { dv[d] = -1; lower = am[d]; dxd = g.uniform ? g.dxv0[d] : a[d] - am[d]; }
        else if (b[d] < pd && pd < bp[d]) { dv[d] = 1; lower = b[d]; dxd = g.uniform ? g.dxv0[d] : bp[d] - b[d]; }
        else { near = false; lower = a[d]; dxd = dx0; }
        dest_ok = dest_ok & (pd < lower + dxd);
    }
    if (in) return JP_CLS_STAY;
    if (!indom) return JP_CODE_DELETE;
    if (!near) return JP_CLS_CPLX + 1;
    if (dv[0] == 0 && dv[1] == 0 && dv[2] == 0) return JP_CLS_CPLX + 2;
    if (!dest_ok) return JP_CLS_CPLX + 3;
    return (dv[0] + 1) + 3 * (dv[1] + 1) + (N == 3 ? 9 * (dv[2] + 1) : 9);
}

// Shortcut for jp_classify_particle on range grids (g.cls_fast): u = (p - a) / dx in single precision
// locates the particle relative to its storage cell; when every coordinate is at least JP_CLS_EPS cell
// widths away from a vertex, each strict comparison of jp_classify_particle is decided with a margin
// (1e-4 dx) that is orders of magnitude above every rounding involved (u: < 2e-7; vertices vs. the affine
// model: <= 1e-6 dx, checked by jp_grid_build; fl(a + dx), fl(lower + dx): 1 ulp), so the result is THE
// SAME code.  Otherwise (within 1e-4 dx of a vertex, more than one cell away, NaN / Inf) it returns -1
// and the caller evaluates jp_classify_particle.  ci = storage cell, a = its lower vertices.
#define JP_CLS_EPS 1.0e-4f
template <int N>
JP_HD int jp_classify_fast(const JpGrid &g, const int *ci, const double *a, const double *p) {
    bool sure = true, del = false;
    int code = 13;                                         // direction (0, 0, 0); 2-D codes carry dz = 0
#pragma unroll
    for (int d = 0; d < N; d++) {
        const float u = (float)((p[d] - a[d]) * g.inv_dv[d]);
        const float kf = floorf(u);                        // -1 / 0 / +1: left neighbour / own cell / right neighbour
        const float fr = u - kf;
        const bool ok = fr > JP_CLS_EPS && fr < 1.0f - JP_CLS_EPS && fabsf(kf) <= 1.0f;   // false for NaN / Inf
        sure = sure && ok;
        const int dv = ok ? (int)kf : 0;
        del = del || (unsigned)(ci[d] + dv) >= (unsigned)g.n[d];   // the neighbour does not exist: p is outside the domain
        code += dv * (d == 0 ? 1 : d == 1 ? 3 : 9);
    }
    if (!sure) return -1;
    return code == 13 ? JP_CLS_STAY : del ? JP_CODE_DELETE : code;
}

// isincell (src/Particles/utils.jl:7-15): strict, upper edge = fl(xv + dx)
template <int N> JP_HD bool jp_isincell(const double *p, const double *corner, const double *dx) {
    bool in = true;
    for (int d = 0; d < N; d++) in = in & (corner[d] < p[d]) & (p[d] < corner[d] + dx[d]);
    return in;
}

// distance (src/Interpolations/utils.jl:9-19): sqrt(((a1-b1)^2 + (a2-b2)^2) + (a3-b3)^2)
template <int N> JP_HD double jp_distance(const double *a, const double *b) {
    double s = (a[0] - b[0]) * (a[0] - b[0]);
    for (int d = 1; d < N; d++) s = s + (a[d] - b[d]) * (a[d] - b[d]);
    return sqrt(s);
}

// bilinear_weight (src/Interpolations/particle_to_grid.jl:211-223,
// src/PhaseRatios/utils.jl:64-74): prod_d muladd(-|a-b|, inv(d), 1); idi = inv(di)
template <int N> JP_HD double jp_bilinear_weight(const double *a, const double *b, const double *idi) {
    double val = 1.0;
    for (int d = 0; d < N; d++) val *= fma(-fabs(a[d] - b[d]), idi[d], 1.0);
    return val;
}

// ---------------------------------------------------------------------------
// init_particles: fill_coords_index! (src/Particles/particles_utils.jl:168-194)
// for one cell; dead slots get NaN / 0.
template <int N>
JP_HD void jp_init_cell(const JpGrid &g, double *const *coords, uint8_t *index, int npq, uint64_t seed, int64_t c) {
    const int NQ = N == 2 ? 4 : 8;
    int ci[3];
    jp_cell_ijk<N>(g, c, ci);
    double x0[3], dx[3];
    for (int d = 0; d < N; d++) { x0[d] = g.xv[d][ci[d]]; dx[d] = jp_d_of(g.xv[d], g.uniform, ci[d]); }
    const int nlive = npq * NQ;
    for (int l = 0; l < g.S; l++) {
        const int64_t e = c + (int64_t)l * g.C;
        if (l < nlive) {
            const int iq = l / npq;
            double r[3];
            jp_rand3(seed, 0u, 0u, (uint32_t)c, (uint32_t)l, r);
            for (int d = 0; d < N; d++) {
                const double xq = x0[d] + dx[d] * (double)((iq >> d) & 1) / 2;
                coords[d][e] = xq + dx[d] / 2 * r[d];
            }
            index[e] = 1;
        } else {
            for (int d = 0; d < N; d++) coords[d][e] = NAN;
            index[e] = 0;
        }
    }
}

// ---------------------------------------------------------------------------
// move_particles!, split into an order-free classification and the ordered
// 3^N colour sweeps working on per-cell occupancy words.
//
// occ[c]   bit s = slot s of cell c is live
// leave[c] bit s = slot s must be visited by the sweep of cell c, i.e. the
//                  particle fails the strict isincell test of the cell it is
//                  stored in (src/Particles/move_safe.jl:91).
// Classification of one (cell, slot): returns true when the particle leaves.
template <int N>
JP_HD bool jp_move_leaves(const JpGrid &g, const int *ci, const double *p) {
    double corner[3], dx[3];
    for (int d = 0; d < N; d++) { corner[d] = g.xv[d][ci[d]]; dx[d] = jp_d_of(g.xv[d], g.uniform, ci[d]); }
    return !jp_isincell<N>(p, corner, dx);
}

// One source cell of one colour sweep: literal move_kernel!
// (src/Particles/move_safe.jl:72-125) restricted to the slots flagged in
// leave[c]; free-slot search (find_free_memory :192-197) on occupancy words.
// A particle placed into a cell whose strict isincell test it fails is flagged
// in that cell's leave word so that the cell's own (later) sweep revisits it,
// exactly as the reference's slot loop would.
// stats: [0] moved, [1] dropped, [2] deleted (accumulated by the caller).
template <int N>
JP_HD void jp_move_cell(const JpGrid &g, double *const *coords, uint8_t *index, const JpArgs &args,
                        uint64_t *occ, uint64_t *leave, int64_t c, const int *ci, int *stats, int cursor = 0, bool compact = false) {
    uint64_t lv = leave[c];
    if (lv == 0) return;
    const int S = g.S;
    const uint64_t smask = S == 64 ? ~0ull : ((1ull << S) - 1);
    double lo[3], hi[3];
    for (int d = 0; d < N; d++) { lo[d] = g.xv[d][0]; hi[d] = g.xv[d][g.n[d]]; }
    uint64_t occ_c = occ[c];
    while (lv) {
#if defined(__CUDA_ARCH__)
        const int ip = __ffsll((long long)lv) - 1;
#else
        const int ip = __builtin_ctzll(lv);
#endif
        lv &= lv - 1;
        const int64_t e = c + (int64_t)ip * g.C;
        double p[3];
        for (int d = 0; d < N; d++) p[d] = coords[d][e];
        bool indom = true;
        for (int d = 0; d < N; d++) indom = indom && (lo[d] < p[d] && p[d] < hi[d]);
        // vacate the source slot (both the delete and the move branch do)
        double cache[JP_MAX_ARGS];
        for (int a = 0; a < args.n; a++) { cache[a] = args.a[a][e]; args.a[a][e] = NAN; }
        for (int d = 0; d < N; d++) coords[d][e] = NAN;
        index[e] = 0;
        occ_c &= ~(1ull << ip);
        if (!indom) { stats[2]++; continue; }
        int nc[3] = {0, 0, 0};
        for (int d = 0; d < N; d++) nc[d] = jp_bisect1(p[d], g.xv[d], g.n[d] + 1, ci[d] + 1) - 1;
        const int64_t c2 = jp_cell_lin<N>(g, nc);
        const bool same = c2 == c;
        uint64_t o2 = same ? occ_c : occ[c2];
        const uint64_t freebits = ~o2 & smask & (cursor >= 64 ? 0ull : (~0ull << cursor));
        if (freebits == 0) { stats[1]++; continue; }
#if defined(__CUDA_ARCH__)
        const int fs = __ffsll((long long)freebits) - 1;
#else
        const int fs = __builtin_ctzll(freebits);
#endif
        if (!compact) cursor = fs;
        o2 |= 1ull << fs;
        const int64_t e2 = c2 + (int64_t)fs * g.C;
        index[e2] = 1;
        for (int d = 0; d < N; d++) coords[d][e2] = p[d];
        for (int a = 0; a < args.n; a++) args.a[a][e2] = cache[a];
        stats[0]++;
        // does it pass the strict isincell test of its new cell?
        double corner[3], dx[3];
        for (int d = 0; d < N; d++) { corner[d] = g.xv[d][nc[d]]; dx[d] = jp_d_of(g.xv[d], g.uniform, nc[d]); }
        const bool fails = !jp_isincell<N>(p, corner, dx);
        if (same) {
            occ_c = o2;
            if (fails && fs > ip) lv |= 1ull << fs;      // revisited later in this very loop
        } else {
            occ[c2] = o2;
            if (fails) leave[c2] |= 1ull << fs;          // visited if that cell's sweep is still to come
        }
    }
    occ[c] = occ_c;
    leave[c] = 0;
}

// max_xcell > JP_MAX_SLOTS ("wide" cells; the reference's own tests use max_xcell = 80 and 150,
// test/test_2D.jl:437, test/test_3D.jl:369): one source cell of one colour sweep as the literal slot loop of
// move_kernel! (src/Particles/move_safe.jl:72-125) on the index bytes themselves -- no occupancy words.
// A particle re-slotted into its own cell at a higher slot, or into a cell whose sweep is still to come, is
// met again by the loop exactly as in the reference.
template <int N>
JP_HD void jp_move_cell_wide(const JpGrid &g, double *const *coords, uint8_t *index, const JpArgs &args,
                             int64_t c, const int *ci, int *stats, bool compact) {
    const int S = g.S;
    const int64_t C = g.C;
    double lo[3], hi[3], corner[3], dx[3];
    for (int d = 0; d < N; d++) {
        lo[d] = g.xv[d][0]; hi[d] = g.xv[d][g.n[d]];
        corner[d] = g.xv[d][ci[d]]; dx[d] = jp_d_of(g.xv[d], g.uniform, ci[d]);
    }
    int cursor = 0;                                   // starting_point - 1
    for (int ip = 0; ip < S; ip++) {
        const int64_t e = c + (int64_t)ip * C;
        if (!index[e]) continue;
        double p[3];
        for (int d = 0; d < N; d++) p[d] = coords[d][e];
        if (jp_isincell<N>(p, corner, dx)) continue;
        bool indom = true;
        for (int d = 0; d < N; d++) indom = indom && (lo[d] < p[d] && p[d] < hi[d]);
        double cache[JP_MAX_ARGS];
        for (int a = 0; a < args.n; a++) { cache[a] = args.a[a][e]; args.a[a][e] = NAN; }
        for (int d = 0; d < N; d++) coords[d][e] = NAN;
        index[e] = 0;
        if (!indom) { stats[2]++; continue; }
        int nc[3] = {0, 0, 0};
        for (int d = 0; d < N; d++) nc[d] = jp_bisect1(p[d], g.xv[d], g.n[d] + 1, ci[d] + 1) - 1;
        const int64_t c2 = jp_cell_lin<N>(g, nc);
        int fs = -1;                                  // find_free_memory(starting_point, index, new_cell)
        for (int i = cursor; i < S; i++)
            if (!index[c2 + (int64_t)i * C]) { fs = i; break; }
        if (fs < 0) { stats[1]++; continue; }
        if (!compact) cursor = fs;
        const int64_t e2 = c2 + (int64_t)fs * C;
        index[e2] = 1;
        for (int d = 0; d < N; d++) coords[d][e2] = p[d];
        for (int a = 0; a < args.n; a++) args.a[a][e2] = cache[a];
        stats[0]++;
    }
}

// clean_particles! (src/Particles/move_safe.jl:289-320) for one live slot:
// true = particle must be removed.
template <int N>
JP_HD bool jp_clean_removes(const JpGrid &g, const int *ci, const double *p) { return jp_move_leaves<N>(g, ci, p); }

// ---------------------------------------------------------------------------
// inject_particles! for one cell: literal _inject_particles!
// (src/Particles/injection.jl:68-131) with index_min_distance (:330-393),
// new_particle (:411-417), quadrant_corners (:442-461).
// Returns the number of injected particles.
template <int N>
JP_HD int jp_inject_cell(const JpGrid &g, double *const *coords, uint8_t *index, const JpArgs &args,
                         int min_xcell, uint64_t seed, uint32_t step, int64_t c, const int *ci) {
    const int S = g.S, NQ = N == 2 ? 4 : 8;
    const int64_t C = g.C;
    const int nx = g.n[0], ny = g.n[1], nz = N == 3 ? g.n[2] : 1;
    double xvc[3], dq[3];
    for (int d = 0; d < N; d++) { xvc[d] = g.xv[d][ci[d]]; dq[d] = jp_d_of(g.xv[d], g.uniform, ci[d]) / 2; }
    const int min_xq = (min_xcell + NQ - 1) / NQ;
    int injected = 0;
    for (int iq = 0; iq < NQ; iq++) {
        double vq[3];
        for (int d = 0; d < N; d++) vq[d] = xvc[d] + dq[d] * (double)((iq >> d) & 1);
        int num = 0;
        for (int i = 0; i < S; i++) {
            const int64_t e = c + (int64_t)i * C;
            if (!index[e]) continue;
            double p[3];
            for (int d = 0; d < N; d++) p[d] = coords[d][e];
            num += jp_isincell<N>(p, vq, dq) ? 1 : 0;
        }
        if (num >= min_xq) break;
        for (int i = 0; i < S; i++) {
            const int64_t e = c + (int64_t)i * C;
            if (index[e]) continue;
            num++;
            double r[3], pn[3];
            jp_rand3(seed, 1u, step, (uint32_t)c, (uint32_t)i, r);
            for (int d = 0; d < N; d++) pn[d] = vq[d] + dq[d] * fma(0.95, r[d], 0.05);
            for (int d = 0; d < N; d++) coords[d][e] = pn[d];
            index[e] = 1;
            injected++;
            double dmin = INFINITY;
            int64_t emin = -1;
            for (int kk = (N == 3 ? ci[2] - 1 : 0); kk <= (N == 3 ? ci[2] + 1 : 0); kk++)
                for (int jj = ci[1] - 1; jj <= ci[1] + 1; jj++)
                    for (int ii = ci[0] - 1; ii <= ci[0] + 1; ii++) {
                        if (ii < 0 || jj < 0 || kk < 0 || ii >= nx || jj >= ny || kk >= nz) continue;
                        const int64_t c2 = ii + (int64_t)nx * (jj + (int64_t)ny * kk);
                        for (int ip = 0; ip < S; ip++) {
                            if (c2 == c && ip == i) continue;
                            const int64_t e2 = c2 + (int64_t)ip * C;
                            if (!index[e2]) continue;
                            double q[3];
                            for (int d = 0; d < N; d++) q[d] = coords[d][e2];
                            const double dist = jp_distance<N>(q, pn);
                            if (dist < dmin) { dmin = dist; emin = e2; }
                        }
                    }
            if (emin >= 0)
                for (int a = 0; a < args.n; a++) args.a[a][e] = args.a[a][emin];
            if (num >= min_xq) break;
        }
    }
    return injected;
}

// Would the reference inject anything into this cell?  The quadrant loop
// `break`s as soon as one quadrant holds >= min_xq particles
// (src/Particles/injection.jl:100), so injection happens iff the FIRST
// quadrant is deficient and the cell has a free slot.  nq0 = live particles
// strictly inside quadrant 0, nlive = live slots.
JP_HD bool jp_inject_candidate(int nq0, int nlive, int S, int min_xq) { return nq0 < min_xq && nlive < S; }

// ---------------------------------------------------------------------------
// cell-local interpolation bodies
// grid2particle! (src/Interpolations/grid_to_particle.jl:64-82, :265-273)
template <int N>
JP_HD double jp_g2p(const double *v, const double *xcorner, const double *idx, const double *p) {
    double t[3];
    for (int d = 0; d < N; d++) t[d] = (p[d] - xcorner[d]) * idx[d];
    return jp_lerp<N>(v, t);
}

// centroid2particle! (src/Interpolations/centroid_to_particle.jl:29-46, :73-76)
template <int N>
JP_HD double jp_c2p(const JpGrid &g, const double *Fc, const int *ci, const double *p) {
    int cc[3] = {0, 0, 0};
    double t[3];
    for (int d = 0; d < N; d++) {
        int i1 = ci[d] + 1;
        if (p[d] < g.xc[d][ci[d]]) i1 -= 1;
        const int hi1 = g.n[d] - 1;
        if (i1 > hi1) i1 = hi1; else if (i1 < 1) i1 = 1;
        cc[d] = i1 - 1;
        const double dx = jp_d_of(g.xc[d], g.uniform, cc[d]);
        t[d] = (p[d] - g.xc[d][cc[d]]) * (1.0 / dx);
    }
    const int64_t s1 = g.n[0], s2 = (int64_t)g.n[0] * g.n[1];
    const int64_t b = cc[0] + s1 * cc[1] + (N == 3 ? s2 * cc[2] : 0);
    double v[8];
    jp_corners<N>(Fc, b, s1, s2, v);
    return jp_lerp<N>(v, t);
}

// particle2grid! weight (distance_weight order=2, src/Interpolations/particle_to_grid.jl:203-205)
template <int N> JP_HD double jp_p2g_weight(const double *xnode, const double *p) {
    const double dist = jp_distance<N>(xnode, p);
    return 1.0 / (dist * dist);
}
