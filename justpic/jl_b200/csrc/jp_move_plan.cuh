// jp_move_plan.cuh -- move_particles! as  classify -> plan -> gather -> scatter.
//
// The reference's 3^N ordered colour sweeps (src/Particles/move_safe.jl:21-125)
// decide, for every particle that left its cell, WHICH slot of WHICH cell it
// lands in.  That decision needs only per-cell occupancy bit-masks and the
// destination of each leaving particle -- not the particle payload.  So:
//
//  A. k_move_classify2 (one coalesced pass over coords + mask): per cell the
//     occupancy word, the leave word and a packed list of 5-bit destination
//     codes (one of the 3^N neighbours, or "left the domain").  Anything the
//     planner cannot express exactly -- a particle that fails isincell but
//     bisects back into its own cell or fails isincell in its destination (on a
//     face / in the fl(x+dx) ulp gap) or a displacement of more than one cell --
//     raises a flag and the whole call takes the direct sweep kernels
//     (k_move_sweep), which handle every case.
//  B. k_move_plan x 3^N (ordered, same colour order as the reference; touches
//     8-byte words only): literal slot logic -- vacate in slot order, first free
//     slot >= cursor, cursor shared across destinations, drop when full -- on the
//     occupancy words; the slot given to each leaver is recorded (7 bits each).
//  C. k_move_finalize + exclusive scan: arrival mask (= final occupancy minus the
//     original occupants that stayed) / count per cell -> offsets into a compact
//     staging buffer (ordered by destination cell, then slot).
//  D. k_move_gather (source-centric, slot-synchronous, streaming): every placed
//     leaver's payload (coords + fields) -> staging.
//  E. k_move_scatter: one coalesced slot-synchronous pass writes arrivals from
//     staging, NaN into vacated slots that stayed empty, and the mask bytes;
//     x-adjacent lanes hit the same 32-byte sectors in the same instruction.
//
// Payload bytes thus move through a few streaming passes instead of ~20 random
// read-modify-write sector transactions per migrant.  Slot assignment is
// bit-identical to the direct sweeps by construction.
#pragma once
#include "jp_core.h"

#define JP_CODE_DELETE 27

// struct MovePlanWs is defined in justpic_sm100a.cu (it is a member of jp_ctx)

__device__ __forceinline__ int jp_dir_code(const int *dv, int N) {
    return (dv[0] + 1) + 3 * (dv[1] + 1) + (N == 3 ? 9 * (dv[2] + 1) : 9);
}
__device__ __forceinline__ void jp_code_dir(int code, int *dv) {
    dv[0] = code % 3 - 1; dv[1] = (code / 3) % 3 - 1; dv[2] = code / 9 - 1;
}

// ---- A. classify
// Destination by comparisons against the four vertices around the storage cell:
// a particle strictly inside (xv[j], xv[j+1]) for j in {i-1, i, i+1} is exactly where the
// reference's seeded bisection puts it; anything else (on a vertex, further away) is
// left to the direct sweeps.
template <int N>
__global__ void __launch_bounds__(256, 3) k_move_classify2(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, MovePlanWs ws,
                                                        unsigned int *complex_flag) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    uint64_t lv = 0, codew = 0;
    int nl = 0;
    unsigned cplx = 0;     // reason bits: 1 far, 2 same cell (tie), 4 fails isincell in destination
    double am[3], a[3], b[3], bp[3], dx[3], lo[3], hi[3];
    if (ok)
        for (int d = 0; d < N; d++) {
            const double *xv = g.xv[d];
            const int i = ci[d];
            a[d] = xv[i]; b[d] = xv[i + 1];
            am[d] = i > 0 ? xv[i - 1] : NAN;
            bp[d] = i + 2 <= g.n[d] ? xv[i + 2] : NAN;
            dx[d] = jp_d_of(xv, g.uniform, i);
            lo[d] = xv[0]; hi[d] = xv[g.n[d]];
        }
    for (int s = 0; s < g.S; s++) {
        const bool live = (m >> s) & 1ull;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (live) {
            const int64_t e = c + (int64_t)s * g.C;
            double p[3];
#pragma unroll
            for (int d = 0; d < N; d++) p[d] = co.p[d][e];
            bool incell = true;
#pragma unroll
            for (int d = 0; d < N; d++) incell = incell & (a[d] < p[d]) & (p[d] < a[d] + dx[d]);
            if (!incell) {
                lv |= 1ull << s;
                bool indom = true;
#pragma unroll
                for (int d = 0; d < N; d++) indom = indom && (lo[d] < p[d] && p[d] < hi[d]);
                int code = JP_CODE_DELETE;
                if (indom) {
                    int dv[3] = {0, 0, 0};
                    bool far = false, dest_ok = true;
#pragma unroll
                    for (int d = 0; d < N; d++) {
                        const double pd = p[d];
                        double lower, dxd;
                        if (a[d] < pd && pd < b[d]) { dv[d] = 0; lower = a[d]; dxd = g.uniform ? dx[d] : b[d] - a[d]; }
                        else if (am[d] < pd && pd < a[d]) { dv[d] = -1; lower = am[d]; dxd = g.uniform ? dx[d] : a[d] - am[d]; }
                        else if (b[d] < pd && pd < bp[d]) { dv[d] = 1; lower = b[d]; dxd = g.uniform ? dx[d] : bp[d] - b[d]; }
                        else { far = true; lower = a[d]; dxd = dx[d]; }        // on a vertex or more than one cell away
                        dest_ok = dest_ok && (lower < pd) && (pd < lower + dxd);
                    }
                    const bool same = dv[0] == 0 && dv[1] == 0 && (N == 2 || dv[2] == 0);
                    if (far || same || !dest_ok) cplx |= (far ? 1u : 0u) | ((same && !far) ? 2u : 0u) | ((!dest_ok && !far) ? 4u : 0u);
                    else code = jp_dir_code(dv, N);
                }
                codew |= (uint64_t)code << (5 * (nl % 12));
                nl++;
                if (nl % 12 == 0) { ws.code[(int64_t)(nl / 12 - 1) * g.C + c] = codew; codew = 0; }
            }
        }
    }
    if (ok) {
        ws.occ[c] = m; ws.occ0[c] = m; ws.leave[c] = lv;
        if (nl % 12 != 0) ws.code[(int64_t)(nl / 12) * g.C + c] = codew;
    }
    const unsigned wc = __reduce_or_sync(0xffffffffu, cplx);
    if (wc && threadIdx.x == 0) atomicOr(complex_flag, wc);
}

// ---- B. one colour of the plan (thread = source cell; 8-byte words only).
// Literal slot logic of move_kernel! (src/Particles/move_safe.jl:72-125) on the occupancy
// words; the slot given to the k-th leaver goes to res (7 bits: slot | placed << 6).
template <int N>
__global__ void __launch_bounds__(256) k_move_plan(JpGrid g, MovePlanWs ws, int ox, int oy, int oz, int ncx, int ncy, int64_t ncol,
                                                   long long *stats) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncol) return;
    int ci[3];
    ci[0] = 3 * (int)(t % ncx) + ox;
    ci[1] = 3 * (int)((t / ncx) % ncy) + oy;
    ci[2] = N == 3 ? 3 * (int)(t / ((int64_t)ncx * ncy)) + oz : 0;
    if (ci[0] >= g.n[0] || ci[1] >= g.n[1] || (N == 3 && ci[2] >= g.n[2])) return;
    const int64_t c = jp_cell_lin<N>(g, ci);
    uint64_t lv = ws.leave[c];
    if (lv == 0) return;
    const uint64_t smask = g.S == 64 ? ~0ull : ((1ull << g.S) - 1);
    uint64_t occ_c = ws.occ[c];
    uint64_t codew = 0, resw = 0;
    int cursor = 0, k = 0, n_dropped = 0, n_deleted = 0;
    while (lv) {
        const int ip = __ffsll((long long)lv) - 1;
        lv &= lv - 1;
        if (k % 12 == 0) codew = ws.code[(int64_t)(k / 12) * g.C + c];
        if (k % 9 == 0 && k > 0) { ws.res[(int64_t)(k / 9 - 1) * g.C + c] = resw; resw = 0; }
        const int code = (int)((codew >> (5 * (k % 12))) & 31);
        const int kk = k++;
        occ_c &= ~(1ull << ip);
        if (code == JP_CODE_DELETE) { n_deleted++; continue; }
        int dv[3];
        jp_code_dir(code, dv);
        const int64_t c2 = c + dv[0] + (int64_t)g.n[0] * (dv[1] + (N == 3 ? (int64_t)g.n[1] * dv[2] : 0));
        const uint64_t o2 = ws.occ[c2];
        const uint64_t freebits = ~o2 & smask & (~0ull << cursor);
        if (freebits == 0) { n_dropped++; continue; }
        const int fs = __ffsll((long long)freebits) - 1;
        cursor = fs;
        ws.occ[c2] = o2 | (1ull << fs);
        resw |= (uint64_t)(fs | 64) << (7 * (kk % 9));
    }
    ws.occ[c] = occ_c;
    ws.res[(int64_t)((k - 1) / 9) * g.C + c] = resw;
    if (n_dropped) atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_dropped);
    if (n_deleted) atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_deleted);
}

// ---- C. arrival mask + count per cell: every slot occupied at the end that is not a
// non-leaving original occupant holds an arrival.
template <int N>
__global__ void __launch_bounds__(256) k_move_finalize(JpGrid g, MovePlanWs ws) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.C) return;
    const uint64_t am = ws.occ[c] & (~ws.occ0[c] | ws.leave[c]);
    ws.arrmask[c] = am;
    ws.cnt[c] = (uint32_t)__popcll(am);
}

struct MoveArrays { double *a[JP_MAX_ARGS + 3]; int n; };

// ---- D. gather (source-centric, slot-synchronous => every source sector is read once, in
// streaming order): payload of each placed leaver -> staging[off[dest] + rank of its slot].
// Slots are handled in batches of U with all loads of a batch issued before the stores.
#define JP_MV_U 4
#define JP_MV_A 4      // arrays handled per register batch (coords + fields); more arrays loop again
template <int N>
__global__ void __launch_bounds__(256, 3) k_move_gather(JpGrid g, MovePlanWs ws, MoveArrays arrs, double *__restrict__ stage, int64_t M /* staging stride */) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t lv = ok ? ws.leave[c] : 0;
    if (!__any_sync(0xffffffffu, lv != 0)) return;
    uint64_t codew = 0, resw = 0;
    int codepl = -1, respl = -1;
    for (int s0 = 0; s0 < g.S; s0 += JP_MV_U) {
        const unsigned bits = (unsigned)(lv >> s0) & ((1u << JP_MV_U) - 1u);
        if (!__any_sync(0xffffffffu, bits != 0)) continue;
        int64_t pos[JP_MV_U], e[JP_MV_U];
        bool act[JP_MV_U];
#pragma unroll
        for (int u = 0; u < JP_MV_U; u++) {
            const int s = s0 + u;
            act[u] = false; pos[u] = 0; e[u] = c + (int64_t)s * g.C;
            if ((bits >> u) & 1u) {
                const int k = __popcll(lv & ((1ull << s) - 1));
                if (k / 9 != respl) { respl = k / 9; resw = ws.res[(int64_t)respl * g.C + c]; }
                const int r = (int)((resw >> (7 * (k % 9))) & 127);
                if (r & 64) {
                    const int fs = r & 63;
                    if (k / 12 != codepl) { codepl = k / 12; codew = ws.code[(int64_t)codepl * g.C + c]; }
                    const int code = (int)((codew >> (5 * (k % 12))) & 31);
                    int dv[3];
                    jp_code_dir(code, dv);
                    const int64_t c2 = c + dv[0] + (int64_t)g.n[0] * (dv[1] + (N == 3 ? (int64_t)g.n[1] * dv[2] : 0));
                    pos[u] = (int64_t)ws.off[c2] + __popcll(ws.arrmask[c2] & ((1ull << fs) - 1));
                    act[u] = true;
                }
            }
        }
        for (int a0 = 0; a0 < arrs.n; a0 += JP_MV_A) {
            double v[JP_MV_U][JP_MV_A];
#pragma unroll
            for (int u = 0; u < JP_MV_U; u++)
#pragma unroll
                for (int a = 0; a < JP_MV_A; a++)
                    if (act[u] && a0 + a < arrs.n) v[u][a] = arrs.a[a0 + a][e[u]];
#pragma unroll
            for (int u = 0; u < JP_MV_U; u++)
#pragma unroll
                for (int a = 0; a < JP_MV_A; a++)
                    if (act[u] && a0 + a < arrs.n) stage[(int64_t)(a0 + a) * M + pos[u]] = v[u][a];
        }
    }
}

// ---- E. scatter: arrivals from staging, NaN into vacated slots, mask bytes
template <int N>
__global__ void __launch_bounds__(256, 3) k_move_scatter(JpGrid g, MovePlanWs ws, MoveArrays arrs, uint8_t *index, const double *__restrict__ stage, int64_t M) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t amask = ok ? ws.arrmask[c] : 0, lmask = ok ? ws.leave[c] : 0;
    const uint64_t changed = amask | lmask;
    if (!__any_sync(0xffffffffu, changed != 0)) return;
    const int64_t base = ok ? ws.off[c] : 0;
    for (int s0 = 0; s0 < g.S; s0 += JP_MV_U) {
        const unsigned chb = (unsigned)(changed >> s0) & ((1u << JP_MV_U) - 1u);
        if (!__any_sync(0xffffffffu, chb != 0)) continue;
        const unsigned arb = (unsigned)(amask >> s0) & ((1u << JP_MV_U) - 1u);
        int64_t pos[JP_MV_U];
#pragma unroll
        for (int u = 0; u < JP_MV_U; u++) pos[u] = base + __popcll(amask & ((1ull << (s0 + u)) - 1));
        for (int a0 = 0; a0 < arrs.n; a0 += JP_MV_A) {
            double v[JP_MV_U][JP_MV_A];
#pragma unroll
            for (int u = 0; u < JP_MV_U; u++)
#pragma unroll
                for (int a = 0; a < JP_MV_A; a++)
                    if (a0 + a < arrs.n && ((chb >> u) & 1u)) v[u][a] = ((arb >> u) & 1u) ? stage[(int64_t)(a0 + a) * M + pos[u]] : NAN;
#pragma unroll
            for (int u = 0; u < JP_MV_U; u++)
#pragma unroll
                for (int a = 0; a < JP_MV_A; a++)
                    if (a0 + a < arrs.n && ((chb >> u) & 1u)) arrs.a[a0 + a][c + (int64_t)(s0 + u) * g.C] = v[u][a];
        }
#pragma unroll
        for (int u = 0; u < JP_MV_U; u++) {
            if ((chb >> u) & 1u) {
                const int s = s0 + u;
                if ((arb >> u) & 1u) { if (!((lmask >> s) & 1ull)) index[c + (int64_t)s * g.C] = 1; }
                else index[c + (int64_t)s * g.C] = 0;
            }
        }
    }
}

__global__ void k_move_set_moved(long long *stats, const uint32_t *total) { stats[0] = (long long)*total; }
