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

// ---- faces (velocity nodes): thread = cell I -> face I + e_dim (+ the low boundary face when I[dim] == 1)
template <int N, int KMAX>
__global__ void __launch_bounds__(256) k_phase_face(JpGrid g, CPtr3 co, double *__restrict__ ratios, const double *__restrict__ phases, int K, int dim) {
    int ci[3]; int64_t c0;
    if (!tile_cell<N>(g, ci, c0)) return;
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
        jp_phase_acc_cell<N, KMAX, true>(g, co, phases, jp_cell_lin<N>(g, cc), face, di, K, w);
    }
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
__global__ void __launch_bounds__(256) k_phase_midpoint(JpGrid g, CPtr3 co, double *__restrict__ ratios, const double *__restrict__ phases, int K, int plane) {
    int ci[3]; int64_t c0;
    if (!tile_cell<3>(g, ci, c0)) return;
    const int off[3] = {plane != 1, plane != 2, plane != 0};                    // xy (1,1,0), yz (0,1,1), xz (1,0,1)
    const int nm[3] = {g.n[0] + off[0], g.n[1] + off[1], g.n[2] + off[2]};
    const int64_t NM = (int64_t)nm[0] * nm[1] * nm[2];
    double cen[3], mid[3], w[KMAX];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        cen[d] = g.xc[d][ci[d]];
        mid[d] = cen[d] + jp_d_of(g.xv[d], g.uniform, ci[d]) * (double)off[d] / 2;
    }
    jp_midpoint_accumulate<KMAX>(g, co, phases, K, ci, off, mid, w);
    jp_phase_store<KMAX, true>(ratios, (ci[0] + off[0]) + (int64_t)nm[0] * ((ci[1] + off[1]) + (int64_t)nm[1] * (ci[2] + off[2])), NM, K, w);
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
