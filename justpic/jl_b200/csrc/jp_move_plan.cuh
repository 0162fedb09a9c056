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
//     face / in the fl(x+dx) ulp gap), a displacement of more than one cell,
//     more than 24 leavers in one cell -- raises a flag and the whole call takes
//     the direct sweep kernels (k_move_sweep), which handle every case.
//  B. k_move_plan x 3^N (ordered, same colour order as the reference; touches
//     8-byte words only): literal slot logic -- vacate in slot order, first free
//     slot >= cursor, cursor shared across destinations, drop when full -- on the
//     occupancy words; every placement is appended to the destination cell's
//     arrival list (17-bit entries: dest slot, direction, source slot).
//  C. k_move_finalize + exclusive scan: arrival mask / count per cell -> offsets
//     into a compact staging buffer (ordered by destination cell, then slot).
//  D. k_move_gather: every arrival's payload (coords + fields) is read from its
//     source slot (neighbouring cell: L1/L2-local) into staging.
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
#define JP_MAX_PLAN_LEAVERS 24      // 2 words x 12 codes of 5 bits

// struct MovePlanWs is defined in justpic_sm100a.cu (it is a member of jp_ctx)

__device__ __forceinline__ int jp_dir_code(const int *dv, int N) {
    return (dv[0] + 1) + 3 * (dv[1] + 1) + (N == 3 ? 9 * (dv[2] + 1) : 9);
}
__device__ __forceinline__ void jp_code_dir(int code, int *dv) {
    dv[0] = code % 3 - 1; dv[1] = (code / 3) % 3 - 1; dv[2] = code / 9 - 1;
}

// ---- A. classify
template <int N>
__global__ void __launch_bounds__(256) k_move_classify2(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, MovePlanWs ws,
                                                        unsigned int *complex_flag) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    uint64_t lv = 0, code0 = 0, code1 = 0;
    int nl = 0;
    bool cplx = false;
    double corner[3], dx[3], lo[3], hi[3];
    if (ok)
        for (int d = 0; d < N; d++) {
            corner[d] = g.xv[d][ci[d]]; dx[d] = jp_d_of(g.xv[d], g.uniform, ci[d]);
            lo[d] = g.xv[d][0]; hi[d] = g.xv[d][g.n[d]];
        }
    for (int s = 0; s < g.S; s++) {
        const bool live = (m >> s) & 1ull;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (live) {
            const int64_t e = c + (int64_t)s * g.C;
            double p[3];
#pragma unroll
            for (int d = 0; d < N; d++) p[d] = co.p[d][e];
            if (!jp_isincell<N>(p, corner, dx)) {
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
                        const int nc = jp_bisect1(p[d], g.xv[d], g.n[d] + 1, ci[d] + 1) - 1;
                        dv[d] = nc - ci[d];
                        far = far || dv[d] < -1 || dv[d] > 1;
                        const double cd = g.xv[d][nc], dd = jp_d_of(g.xv[d], g.uniform, nc);
                        dest_ok = dest_ok && (cd < p[d]) && (p[d] < cd + dd);
                    }
                    const bool same = dv[0] == 0 && dv[1] == 0 && (N == 2 || dv[2] == 0);
                    if (far || same || !dest_ok) cplx = true;
                    else code = jp_dir_code(dv, N);
                }
                if (nl < 12) code0 |= (uint64_t)code << (5 * nl);
                else if (nl < 24) code1 |= (uint64_t)code << (5 * (nl - 12));
                else cplx = true;
                nl++;
            }
        }
    }
    if (ok) {
        ws.occ[c] = m; ws.leave[c] = lv;
        ws.code[c] = code0; ws.code[g.C + c] = code1;
        ws.arr[c] = 0;                        // plane 0: count = 0, no entries
    }
    if (__any_sync(0xffffffffu, cplx) && threadIdx.x == 0) atomicOr(complex_flag, 1u);
}

// ---- B. one colour of the plan (thread = source cell; words only)
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
    const uint64_t code0 = ws.code[c], code1 = ws.code[g.C + c];
    int cursor = 0, k = 0, n_dropped = 0, n_deleted = 0;
    while (lv) {
        const int ip = __ffsll((long long)lv) - 1;
        lv &= lv - 1;
        const int code = (int)((k < 12 ? code0 >> (5 * k) : code1 >> (5 * (k - 12))) & 31);
        k++;
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
        // append (fs, direction, source slot) to the destination's arrival list
        const uint64_t entry = (uint64_t)fs | ((uint64_t)code << 6) | ((uint64_t)ip << 11);
        uint64_t w0 = ws.arr[c2];
        const int na = (int)((w0 >> 56) & 127);
        const int pl = na / 3, pos = na % 3;
        w0 = (w0 & ~(127ull << 56)) | ((uint64_t)(na + 1) << 56);
        if (pl == 0) ws.arr[c2] = w0 | (entry << (17 * pos));
        else {
            ws.arr[c2] = w0;
            uint64_t *wp = &ws.arr[(int64_t)pl * g.C + c2];
            *wp = pos == 0 ? entry : (*wp | (entry << (17 * pos)));
        }
    }
    ws.occ[c] = occ_c;
    if (n_dropped) atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_dropped);
    if (n_deleted) atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_deleted);
}

__device__ __forceinline__ uint64_t jp_arr_entry(const MovePlanWs &ws, int64_t C, int64_t c, uint64_t w0, int k, uint64_t &wcur, int &plcur) {
    const int pl = k / 3, pos = k % 3;
    if (pl != plcur) { wcur = pl == 0 ? w0 : ws.arr[(int64_t)pl * C + c]; plcur = pl; }
    return (wcur >> (17 * pos)) & 0x1ffffull;
}

// ---- C. arrival mask + count per cell
template <int N>
__global__ void __launch_bounds__(256) k_move_finalize(JpGrid g, MovePlanWs ws) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.C) return;
    const uint64_t w0 = ws.arr[c];
    const int na = (int)((w0 >> 56) & 127);
    uint64_t mask = 0, wcur = w0;
    int plcur = 0;
    for (int k = 0; k < na; k++) mask |= 1ull << (jp_arr_entry(ws, g.C, c, w0, k, wcur, plcur) & 63);
    ws.arrmask[c] = mask;
    ws.cnt[c] = (uint32_t)na;
}

struct MoveArrays { double *a[JP_MAX_ARGS + 3]; int n; };

// ---- D. gather arrivals' payloads into staging (thread = destination cell)
template <int N>
__global__ void __launch_bounds__(256) k_move_gather(JpGrid g, MovePlanWs ws, MoveArrays arrs, double *__restrict__ stage, int64_t M /* staging stride */) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.C) return;
    const uint64_t w0 = ws.arr[c];
    const int na = (int)((w0 >> 56) & 127);
    if (na == 0) return;
    const uint64_t amask = ws.arrmask[c];
    const int64_t base = ws.off[c];
    uint64_t wcur = w0;
    int plcur = 0;
    for (int k = 0; k < na; k++) {
        const uint64_t en = jp_arr_entry(ws, g.C, c, w0, k, wcur, plcur);
        const int fs = (int)(en & 63), code = (int)((en >> 6) & 31), ip = (int)((en >> 11) & 63);
        int dv[3];
        jp_code_dir(code, dv);
        const int64_t csrc = c - (dv[0] + (int64_t)g.n[0] * (dv[1] + (N == 3 ? (int64_t)g.n[1] * dv[2] : 0)));
        const int64_t es = csrc + (int64_t)ip * g.C;
        const int64_t pos = base + __popcll(amask & ((1ull << fs) - 1));
        for (int a = 0; a < arrs.n; a++) stage[(int64_t)a * M + pos] = arrs.a[a][es];
    }
}

// ---- E. scatter: arrivals from staging, NaN into vacated slots, mask bytes
template <int N>
__global__ void __launch_bounds__(256) k_move_scatter(JpGrid g, MovePlanWs ws, MoveArrays arrs, uint8_t *index, const double *__restrict__ stage, int64_t M) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t amask = ok ? ws.arrmask[c] : 0, lmask = ok ? ws.leave[c] : 0;
    const uint64_t changed = amask | lmask;
    if (!__any_sync(0xffffffffu, changed != 0)) return;
    const int64_t base = ok ? ws.off[c] : 0;
    for (int s = 0; s < g.S; s++) {
        const bool ch = (changed >> s) & 1ull;
        if (!__any_sync(0xffffffffu, ch)) continue;
        if (ch) {
            const int64_t e = c + (int64_t)s * g.C;
            if ((amask >> s) & 1ull) {
                const int64_t pos = base + __popcll(amask & ((1ull << s) - 1));
                for (int a = 0; a < arrs.n; a++) arrs.a[a][e] = stage[(int64_t)a * M + pos];
                if (!((lmask >> s) & 1ull)) index[e] = 1;      // was free (else it was live and stays live)
            } else {
                for (int a = 0; a < arrs.n; a++) arrs.a[a][e] = NAN;
                index[e] = 0;
            }
        }
    }
}

__global__ void k_move_set_moved(long long *stats, const uint32_t *total) { stats[0] = (long long)*total; }
