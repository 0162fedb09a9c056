// jp_phase_ratios.cuh -- phase_ratios_vertex! / phase_ratios_face! / phase_ratios_midpoint!
// (src/PhaseRatios/vertices.jl:4-107, midpoints.jl:3-242, utils.jl:64-91).
//
// Literal per-node / per-cell restatements: thread = output node (vertex) or source cell (face,
// midpoint), the contributing cells visited in the reference's order and their slots in slot
// order, so the weighted sums are bit-identical to the reference's accumulation chains.  Each
// output location is written by exactly one thread (see the index analysis in DESIGN.md), so no
// atomics are involved.  Liveness is the reference's any(isnan, p) test on the coordinates.
#pragma once
#include "jp_core.h"

// Accumulate the particles of cell c that lie within half a cell of x.
// CLOSED = false: vertex rule, excluded when |p - x| >= di/2 (vertices.jl:42-48);
// CLOSED = true : isinhalfcell, included when |p - x| <= di/2 (utils.jl:80-81).
template <int N, int KMAX, bool CLOSED>
__device__ __forceinline__ void jp_phase_acc_cell(const JpGrid &g, const CPtr3 &co, const double *__restrict__ phases, int64_t c,
                                                  const double *x, const double *di, int K, double *w) {
    double idi[3], half[3];
#pragma unroll
    for (int d = 0; d < N; d++) { idi[d] = 1.0 / di[d]; half[d] = di[d] / 2; }
    for (int s = 0; s < g.S; s++) {
        const int64_t e = c + (int64_t)s * g.C;
        double p[3];
        bool nan = false, in = true;
#pragma unroll
        for (int d = 0; d < N; d++) { p[d] = co.p[d][e]; nan |= isnan(p[d]); }
        if (nan) continue;
#pragma unroll
        for (int d = 0; d < N; d++) {
            const double a = fabs(p[d] - x[d]);
            in = in && (CLOSED ? a <= half[d] : !(a >= half[d]));
        }
        if (!in) continue;
        const double wt = jp_bilinear_weight<N>(x, p, idi);
        const double ph = phases[e];
#pragma unroll
        for (int k = 0; k < KMAX; k++)
            if (k < K) w[k] = w[k] + (ph == (double)(k + 1) ? wt : copysign(0.0, wt));
    }
}

// w .* inv(sum(w)) and the CellArray store; ZERO_NAN: `w * !isnan(w)` of the face / midpoint kernels
template <int KMAX, bool ZERO_NAN>
__device__ __forceinline__ void jp_phase_store(double *__restrict__ ratios, int64_t node, int64_t NN, int K, const double *w) {
    double sum = w[0];
#pragma unroll
    for (int k = 1; k < KMAX; k++) if (k < K) sum = sum + w[k];
    const double inv = 1.0 / sum;
#pragma unroll
    for (int k = 0; k < KMAX; k++)
        if (k < K) {
            const double v = w[k] * inv;
            ratios[node + (int64_t)k * NN] = (ZERO_NAN && isnan(v)) ? 0.0 : v;
        }
}

// ---- vertices: thread = vertex node, cells I + (-1..0)^N with offset_i (x) outermost
template <int N, int KMAX>
__global__ void __launch_bounds__(256) k_phase_vertex(JpGrid g, CPtr3 co, double *__restrict__ ratios, const double *__restrict__ phases, int K) {
    const int in = blockIdx.x * JP_BX + threadIdx.x, jn = blockIdx.y * JP_BY + threadIdx.y, kn = N == 3 ? blockIdx.z : 0;
    const int nx = g.n[0], ny = g.n[1], nz = N == 3 ? g.n[2] : 1;
    if (in > nx || jn > ny) return;
    const double xv[3] = {g.xv[0][in], g.xv[1][jn], N == 3 ? g.xv[2][kn] : 0.0};
    double w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = 0.0;
    for (int oi = -1; oi <= 0; oi++) {
        const int ic = in + oi;
        if (ic < 0 || ic >= nx) continue;
        for (int oj = -1; oj <= 0; oj++) {
            const int jc = jn + oj;
            if (jc < 0 || jc >= ny) continue;
            for (int ok = (N == 3 ? -1 : 0); ok <= 0; ok++) {
                const int kc = N == 3 ? kn + ok : 0;
                if (N == 3 && (kc < 0 || kc >= nz)) continue;
                const int cc[3] = {ic, jc, kc};
                double di[3];
#pragma unroll
                for (int d = 0; d < N; d++) di[d] = jp_d_of(g.xv[d], g.uniform, cc[d]);
                jp_phase_acc_cell<N, KMAX, false>(g, co, phases, ic + (int64_t)nx * (jc + (int64_t)ny * kc), xv, di, K, w);
            }
        }
    }
    const int64_t NN = (int64_t)(nx + 1) * (ny + 1) * (N == 3 ? nz + 1 : 1);
    jp_phase_store<KMAX, false>(ratios, in + (int64_t)(nx + 1) * (jn + (N == 3 ? (int64_t)(ny + 1) * kn : 0)), NN, K, w);
}

// boundary_only launches (fused mode): threads enumerate only the cells of the plane ci[bdim] == 0 (dense 1-D grid,
// so a warp holds 32 boundary cells instead of one); cells with ci[skipdim] == 0 are left to another launch.
template <int N>
__device__ __forceinline__ bool jp_plane_cell(const JpGrid &g, int bdim, int skipdim, int *ci, int64_t &c) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x * blockDim.y + threadIdx.y * blockDim.x + threadIdx.x;
    const int d1 = bdim == 0 ? 1 : 0, d2 = bdim == 2 ? 1 : 2;                 // the two in-plane dimensions (d2 unused in 2-D)
    const int n1 = g.n[d1], n2 = N == 3 ? g.n[d2] : 1;
    if (t >= (int64_t)n1 * n2) return false;
    ci[0] = ci[1] = ci[2] = 0;
    ci[d1] = (int)(t % n1);
    if (N == 3) ci[d2] = (int)(t / n1);
    if (skipdim >= 0 && ci[skipdim] == 0) return false;
    c = jp_cell_lin<N>(g, ci);
    return true;
}

// ---- faces (velocity nodes): thread = cell I -> face I + e_dim (+ the low boundary face when I[dim] == 1)
template <int N, int KMAX>
__global__ void __launch_bounds__(256) k_phase_face(JpGrid g, CPtr3 co, double *__restrict__ ratios, const double *__restrict__ phases, int K, int dim,
                                                    bool boundary_only) {
    int ci[3]; int64_t c0;
    if (boundary_only ? !jp_plane_cell<N>(g, dim, -1, ci, c0) : !tile_cell<N>(g, ci, c0)) return;   // fused mode: only the low-boundary faces
    const int off[3] = {dim == 0, dim == 1, dim == 2};
    const int nf[3] = {g.n[0] + off[0], g.n[1] + off[1], (N == 3 ? g.n[2] : 1) + (N == 3 ? off[2] : 0)};
    const int64_t NF = (int64_t)nf[0] * nf[1] * nf[2];
    double di[3], cen[3], face[3], w[KMAX];
#pragma unroll
    for (int d = 0; d < N; d++) {
        di[d] = jp_d_of(g.xv[d], g.uniform, ci[d]);
        cen[d] = g.xc[d][ci[d]];
        face[d] = cen[d] + di[d] * (double)off[d] / 2;
    }
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = 0.0;
    for (int pass = 0; pass < 2; pass++) {
        int cc[3] = {0, 0, 0};
#pragma unroll
        for (int d = 0; d < N; d++) { cc[d] = min(ci[d] + (pass ? off[d] : 0), g.n[d] - 1); di[d] = jp_d_of(g.xv[d], g.uniform, cc[d]); }   // `di` is reassigned
        if (!boundary_only) jp_phase_acc_cell<N, KMAX, true>(g, co, phases, jp_cell_lin<N>(g, cc), face, di, K, w);
    }
    if (!boundary_only)
        jp_phase_store<KMAX, true>(ratios, (ci[0] + off[0]) + (int64_t)nf[0] * ((ci[1] + off[1]) + (int64_t)nf[1] * (ci[2] + (N == 3 ? off[2] : 0))), NF, K, w);
    if (ci[dim] == 0) {                                                         // isboundary(offsets, I)
#pragma unroll
        for (int d = 0; d < N; d++) face[d] = cen[d] - di[d] * (double)off[d] / 2;       // di = the last one assigned above
#pragma unroll
        for (int k = 0; k < KMAX; k++) w[k] = 0.0;
        jp_phase_acc_cell<N, KMAX, true>(g, co, phases, c0, face, di, K, w);
        jp_phase_store<KMAX, true>(ratios, ci[0] + (int64_t)nf[0] * (ci[1] + (int64_t)nf[1] * ci[2]), NF, K, w);
    }
}

// ---- edge midpoints (3-D): thread = cell I -> midpoint I + offsets (+ the boundary branch)
template <int KMAX>
__device__ __forceinline__ void jp_midpoint_accumulate(const JpGrid &g, const CPtr3 &co, const double *__restrict__ phases, int K,
                                                       const int *ci, const int *off, const double *mid, double *w) {
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = 0.0;
#pragma unroll 1
    for (int m = 0; m < 4; m++) {                                               // MASK_3D = (1,0,0), (0,1,0), (0,0,1), (1,1,1)
        int cc[3];
        double di[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const int mask = m == 3 ? 1 : (m == d ? 1 : 0);
            cc[d] = min(ci[d] + off[d] * mask, g.n[d] - 1);
            di[d] = jp_d_of(g.xv[d], g.uniform, cc[d]);
        }
        jp_phase_acc_cell<3, KMAX, true>(g, co, phases, jp_cell_lin<3>(g, cc), mid, di, K, w);
    }
}

template <int KMAX>
__global__ void __launch_bounds__(256) k_phase_midpoint(JpGrid g, CPtr3 co, double *__restrict__ ratios, const double *__restrict__ phases, int K, int plane,
                                                        int bdim, int skipdim) {
    // bdim < 0: every cell (literal mode); else only the boundary plane ci[bdim] == 0 minus ci[skipdim] == 0 (fused mode)
    const bool boundary_only = bdim >= 0;
    int ci[3]; int64_t c0;
    if (boundary_only ? !jp_plane_cell<3>(g, bdim, skipdim, ci, c0) : !tile_cell<3>(g, ci, c0)) return;
    const int off[3] = {plane != 1, plane != 2, plane != 0};                    // xy (1,1,0), yz (0,1,1), xz (1,0,1)
    const int nm[3] = {g.n[0] + off[0], g.n[1] + off[1], g.n[2] + off[2]};
    const int64_t NM = (int64_t)nm[0] * nm[1] * nm[2];
    double cen[3], mid[3], w[KMAX];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        cen[d] = g.xc[d][ci[d]];
        mid[d] = cen[d] + jp_d_of(g.xv[d], g.uniform, ci[d]) * (double)off[d] / 2;
    }
    if (!boundary_only) {
        jp_midpoint_accumulate<KMAX>(g, co, phases, K, ci, off, mid, w);
        jp_phase_store<KMAX, true>(ratios, (ci[0] + off[0]) + (int64_t)nm[0] * ((ci[1] + off[1]) + (int64_t)nm[1] * (ci[2] + off[2])), NM, K, w);
    }
    bool boundary = false;
#pragma unroll
    for (int d = 0; d < 3; d++) boundary |= off[d] * (ci[d] + 1) == 1;
    if (boundary) {
        int ob[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
            ob[d] = g.n[d] == off[d] * (ci[d] + 1);                             // lastboundary_offset
            mid[d] = cen[d] - ((jp_d_of(g.xv[d], g.uniform, ci[d]) * (double)off[d]) * (double)(-ob[d])) / 2;
        }
        jp_midpoint_accumulate<KMAX>(g, co, phases, K, ci, off, mid, w);
        for (int pass = 0; pass < 2; pass++)                                    // ((0,0,0), offset_boundary): the `=== false` skip never fires
            jp_phase_store<KMAX, true>(ratios, (ci[0] + pass * ob[0]) + (int64_t)nm[0] * ((ci[1] + pass * ob[1]) + (int64_t)nm[1] * (ci[2] + pass * ob[2])), NM, K, w);
    }
}

// =====================================================================================================
// update_phase_ratios!, FUSED mode (JP_PHASE_FUSED): one pass over the particles for all outputs.
//
// The literal kernels above read every particle 1 (centre) + 8 (vertex) + 3x2 (faces) + 3x4 (midpoints)
// = 27 times per update.  Here pass 1 (thread = cell) reads each particle ONCE and accumulates, per cell,
// its partial weighted sums towards every node it can contribute to: its 2^N corner vertices, its 2 faces
// per dimension and (3-D) its 4 edge midpoints per plane -- at most one of each kind per particle, picked
// by the reference's own half-cell predicates, ties included.  All weights are products of 5 per-dimension
// factors fma(-|p - x|, inv(di), 1), x in {centre, lower face, upper face, lower vertex, upper vertex}, with
// exactly the node coordinates the reference's work-items compute (a face / midpoint coordinate is
// xc[I] + di(I)/2 of the work-item BELOW it, not xv).  Accumulators live in shared memory ([target][phase][thread]).
// Pass 2 (thread = node) adds the partials of the contributing cells in the reference's cell order, upper-
// boundary clamping included (a clamped cell is added twice, as the reference visits it twice), normalises
// and stores.  The low-boundary nodes, which the reference computes in its quirky boundary branches, are
// produced by the literal kernels restricted to the boundary work-items (boundary_only).
// Same terms as the literal kernels, associated as (cell sums) + ...: agrees within the stated 1e-12, not
// bit for bit; the centre ratios are the literal ones.
template <int N> struct PhaseFused {
    static constexpr int NV = 1 << N;                        // vertex targets
    static constexpr int NF = 2 * N;                         // face targets: 2*d + side
    static constexpr int NM = N == 3 ? 12 : 0;               // midpoint targets: 4*plane + u + 2*v
    static constexpr int NT = NV + NF + NM;
    static constexpr int TF = NV, TM = NV + NF;
    static constexpr int THREADS = 128;                      // 32 x 4 cells per CTA
};

template <int N, int KMAX>
__global__ void __launch_bounds__(PhaseFused<N>::THREADS) k_phase_fused_cell(JpGrid g, CPtr3 co, const double *__restrict__ phases, int K,
                                                                             double *__restrict__ center, double *__restrict__ part) {
    using P = PhaseFused<N>;
    extern __shared__ double acc[];                          // [NT][KMAX][THREADS]
    const int tid = threadIdx.y * 32 + threadIdx.x;
    int ci[3];
    ci[0] = blockIdx.x * 32 + threadIdx.x;
    ci[1] = blockIdx.y * 4 + threadIdx.y;
    ci[2] = N == 3 ? blockIdx.z : 0;
    if (ci[0] >= g.n[0] || ci[1] >= g.n[1]) return;
    const int64_t c = jp_cell_lin<N>(g, ci);
    for (int t = 0; t < P::NT * KMAX; t++) acc[t * P::THREADS + tid] = 0.0;
    // per dimension: spacing of this cell and the five reference coordinates
    double idi[3], half[3], xC[3], xFL[3], xFU[3], xVL[3], xVU[3];
#pragma unroll
    for (int d = 0; d < N; d++) {
        const double di = jp_d_of(g.xv[d], g.uniform, ci[d]);
        idi[d] = 1.0 / di; half[d] = di / 2;
        xC[d] = g.xc[d][ci[d]];
        xFU[d] = xC[d] + di * 1.0 / 2;                                                    // work-item I, offsets = 1
        xFL[d] = ci[d] > 0 ? g.xc[d][ci[d] - 1] + jp_d_of(g.xv[d], g.uniform, ci[d] - 1) * 1.0 / 2 : NAN;   // work-item I - e_d (node 0: boundary branch)
        xVL[d] = g.xv[d][ci[d]]; xVU[d] = g.xv[d][ci[d] + 1];
    }
    double wc[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) wc[k] = 0.0;
    // software pipeline over the slots: the next slot's coordinates / phase are in flight while this one is processed
    double pn[3] = {0.0, 0.0, 0.0}, phn = 0.0;
#pragma unroll
    for (int d = 0; d < N; d++) pn[d] = co.p[d][c];
    phn = isnan(pn[0]) ? 0.0 : phases[c];
    for (int s = 0; s < g.S; s++) {
        double p[3];
        bool nan = false;
#pragma unroll
        for (int d = 0; d < N; d++) { p[d] = pn[d]; nan |= isnan(p[d]); }
        const double ph = phn;
        if (s + 1 < g.S) {
            const int64_t en = c + (int64_t)(s + 1) * g.C;
#pragma unroll
            for (int d = 0; d < N; d++) pn[d] = co.p[d][en];
            phn = phases[en];
        }
        if (isnan(p[0])) continue;                               // the centre kernel's liveness test (centers.jl / utils.jl:53)
        double fC[3], fF[3][2], fV[3][2];
        bool pC[3], pF[3][2], pV[3][2];
#pragma unroll
        for (int d = 0; d < N; d++) {
            const double aC = fabs(xC[d] - p[d]), aL = fabs(p[d] - xFL[d]), aU = fabs(p[d] - xFU[d]);
            const double vL = fabs(p[d] - xVL[d]), vU = fabs(p[d] - xVU[d]);
            fC[d] = fma(-aC, idi[d], 1.0); pC[d] = aC <= half[d];
            fF[d][0] = fma(-aL, idi[d], 1.0); pF[d][0] = aL <= half[d];                  // NaN coordinate (node 0) -> false
            fF[d][1] = fma(-aU, idi[d], 1.0); pF[d][1] = aU <= half[d];
            fV[d][0] = fma(-vL, idi[d], 1.0); pV[d][0] = !(vL >= half[d]);
            fV[d][1] = fma(-vU, idi[d], 1.0); pV[d][1] = !(vU >= half[d]);
        }
        // centre (literal: same chain as k_phase)
        {
            double x = 1.0;
#pragma unroll
            for (int d = 0; d < N; d++) x *= fC[d];
#pragma unroll
            for (int k = 0; k < KMAX; k++)
                if (k < K) wc[k] = wc[k] + (ph == (double)(k + 1) ? x : copysign(0.0, x));
        }
        if (nan) continue;                                       // all other kernels: any(isnan, p)
        int kph = -1;
#pragma unroll
        for (int k = 0; k < KMAX; k++) if (k < K && ph == (double)(k + 1)) kph = k;
        if (kph < 0) continue;                                   // phase matches no id: contributes +0 everywhere
        double *a = acc + kph * P::THREADS + tid;
        // vertex: at most one corner (strict predicate)
        for (int q = 0; q < P::NV; q++) {
            bool in = true;
            double w = 1.0;
#pragma unroll
            for (int d = 0; d < N; d++) { const int u = (q >> d) & 1; in = in && pV[d][u]; w *= fV[d][u]; }
            if (in) a[q * KMAX * P::THREADS] += w;
        }
        // faces: dimension d, side u; the other dimensions use the centre coordinate
#pragma unroll
        for (int d = 0; d < N; d++)
            for (int u = 0; u < 2; u++) {
                bool in = pF[d][u];
                double w = 1.0;
#pragma unroll
                for (int dd = 0; dd < N; dd++) { if (dd != d) in = in && pC[dd]; w *= dd == d ? fF[d][u] : fC[dd]; }
                if (in) a[(P::TF + 2 * d + u) * KMAX * P::THREADS] += w;
            }
        // edge midpoints (3-D): plane 0 = xy, 1 = yz, 2 = xz; offset dimensions (da, db), the third uses the centre
        if (N == 3) {
#pragma unroll
            for (int pl = 0; pl < 3; pl++) {
                const int da = pl == 1 ? 1 : 0, db = pl == 0 ? 1 : 2, dc = 3 - da - db;
                for (int uv = 0; uv < 4; uv++) {
                    const int u = uv & 1, v = uv >> 1;
                    if (!(pF[da][u] && pF[db][v] && pC[dc])) continue;
                    double w = 1.0;
#pragma unroll
                    for (int dd = 0; dd < 3; dd++) w *= dd == da ? fF[da][u] : (dd == db ? fF[db][v] : fC[dc]);
                    a[(P::TM + 4 * pl + uv) * KMAX * P::THREADS] += w;
                }
            }
        }
    }
    // centre ratios (final) and the per-cell partials
    jp_phase_store<KMAX, false>(center, c, g.C, K, wc);
    for (int t = 0; t < P::NT; t++)
#pragma unroll
        for (int k = 0; k < KMAX; k++)
            if (k < K) part[((int64_t)t * K + k) * g.C + c] = acc[(t * KMAX + k) * P::THREADS + tid];
}

// pass 2: thread = output node; kind 0 = vertex, 1 = face (dim), 2 = midpoint (plane)
template <int N, int KMAX>
__global__ void __launch_bounds__(256) k_phase_fused_node(JpGrid g, const double *__restrict__ part, double *__restrict__ ratios, int K,
                                                         int kind, int sel) {
    using P = PhaseFused<N>;
    int off[3] = {0, 0, 0};
    if (kind == 0) { off[0] = off[1] = 1; off[2] = N == 3; }
    else if (kind == 1) off[sel] = 1;
    else { off[0] = sel != 1; off[1] = sel != 2; off[2] = sel != 0; }
    const int nn[3] = {g.n[0] + off[0], g.n[1] + off[1], N == 3 ? g.n[2] + off[2] : 1};
    const int nd[3] = {(int)(blockIdx.x * JP_BX + threadIdx.x), (int)(blockIdx.y * JP_BY + threadIdx.y), N == 3 ? (int)blockIdx.z : 0};
    if (nd[0] >= nn[0] || nd[1] >= nn[1]) return;
    double w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = 0.0;
    auto add = [&](const int *cc, int t) {
        const int64_t c = jp_cell_lin<N>(g, cc);
#pragma unroll
        for (int k = 0; k < KMAX; k++)
            if (k < K) w[k] = w[k] + part[((int64_t)t * K + k) * g.C + c];
    };
    if (kind == 0) {                                         // cells V + (-1..0)^N, offset_i outermost (vertices.jl:27-34)
        for (int oi = -1; oi <= 0; oi++)
            for (int oj = -1; oj <= 0; oj++)
                for (int ok = (N == 3 ? -1 : 0); ok <= 0; ok++) {
                    const int cc[3] = {nd[0] + oi, nd[1] + oj, N == 3 ? nd[2] + ok : 0};
                    if (cc[0] < 0 || cc[0] >= g.n[0] || cc[1] < 0 || cc[1] >= g.n[1] || (N == 3 && (cc[2] < 0 || cc[2] >= g.n[2]))) continue;
                    add(cc, (oi < 0 ? 1 : 0) | (oj < 0 ? 2 : 0) | ((N == 3 && ok < 0) ? 4 : 0));
                }
    } else {
        // work-item I' = node - offsets; low-boundary nodes (an offset dimension at index 0) belong to the boundary branch
        int I[3];
        for (int d = 0; d < 3; d++) { I[d] = nd[d] - off[d]; if (d < N && I[d] < 0) return; }
        if (kind == 1) {                                     // cells I' and min(I' + e_d, n - 1) (midpoints.jl:40-42)
            for (int pass = 0; pass < 2; pass++) {
                int cc[3] = {I[0], I[1], I[2]};
                cc[sel] = min(I[sel] + pass, g.n[sel] - 1);
                add(cc, P::TF + 2 * sel + (nd[sel] - cc[sel]));              // side = node - cell (1 when clamped back onto I')
            }
        } else {                                             // MASK_3D order, clamped (midpoints.jl:143-149)
            const int da = sel == 1 ? 1 : 0, db = sel == 0 ? 1 : 2;
            const int mask[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 1}};
            for (int m = 0; m < 4; m++) {
                int cc[3];
                for (int d = 0; d < 3; d++) cc[d] = min(I[d] + off[d] * mask[m][d], g.n[d] - 1);
                add(cc, P::TM + 4 * sel + (nd[da] - cc[da]) + 2 * (nd[db] - cc[db]));
            }
        }
    }
    const int64_t NN = (int64_t)nn[0] * nn[1] * nn[2];
    const int64_t node = nd[0] + (int64_t)nn[0] * (nd[1] + (int64_t)nn[1] * nd[2]);
    if (kind == 0) jp_phase_store<KMAX, false>(ratios, node, NN, K, w);
    else           jp_phase_store<KMAX, true>(ratios, node, NN, K, w);
}
