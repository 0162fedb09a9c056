// jp_move_plan.cuh -- move_particles! as  classify -> plan -> gather -> scatter.
//
// The reference's 3^N ordered colour sweeps (src/Particles/move_safe.jl:21-125)
// decide, for every particle that left its cell, WHICH slot of WHICH cell it
// lands in.  That decision needs only per-cell occupancy bit-masks and the
// destination of each leaving particle -- not the particle payload.  So:
//
//  A. k_move_classify3 (one coalesced pass over coords + mask): per cell the
//     occupancy word, the leave word and a packed list of 5-bit destination
//     codes (one of the 3^N neighbours, or "left the domain").  Anything the
//     planner cannot express exactly -- a particle that fails isincell but
//     bisects back into its own cell or fails isincell in its destination (on a
//     face / in the fl(x+dx) ulp gap) or a displacement of more than one cell --
//     raises a flag and the whole call takes the direct sweep kernels
//     (k_move_sweep_all), which handle every case.
//  B. k_move_plan_all: the 3^N ordered colours (same order as the reference) in one cooperative launch (k_move_plan x 3^N
//     when a colour is larger than what is co-resident); touches
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
#include <cooperative_groups.h>

// JP_CODE_DELETE / JP_CLS_STAY / JP_CLS_CPLX and jp_classify_particle live in jp_core.h

// struct MovePlanWs is defined in justpic_sm100a.cu (it is a member of jp_ctx)

__device__ __forceinline__ int jp_dir_code(const int *dv, int N) {
    return (dv[0] + 1) + 3 * (dv[1] + 1) + (N == 3 ? 9 * (dv[2] + 1) : 9);
}
__device__ __forceinline__ void jp_code_dir(int code, int *dv) {
    dv[0] = code % 3 - 1; dv[1] = (code / 3) % 3 - 1; dv[2] = code / 9 - 1;
}

// ---- A. single-pass classify: every live particle is classified where it is loaded (no second
// look at the leavers, no shared memory): the strict isincell test, the domain test and the
// destination search share their comparisons, and since thread = cell walks its slots in order the
// k-th leaver's code byte is simply shifted into the thread's own code word.  (A first version compacted
// the ~38 % leavers through a shared-memory ring and re-read their coordinates; with 4 CTAs/SM resident
// that second look missed the L2 and cost 40 % extra DRAM traffic at 256^3.)
struct JpBox { int o[3], e[3]; };      // sub-box of cells (origin, extent) for the BOX instantiation
template <int N, bool BOX>
__global__ void __launch_bounds__(256, JP_MINB_CLASSIFY) k_move_classify3(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, MovePlanWs ws,
                                                                         unsigned int *complex_flag, JpBox box) {
    int ci[3]; int64_t c;
    bool ok;
    if (BOX) {
        const int lx = blockIdx.x * JP_BX + threadIdx.x, ly = blockIdx.y * JP_BY + threadIdx.y, lz = N == 3 ? blockIdx.z : 0;
        ci[0] = box.o[0] + lx; ci[1] = box.o[1] + ly; ci[2] = N == 3 ? box.o[2] + lz : 0;
        ok = lx < box.e[0] && ly < box.e[1] && ci[0] < g.n[0] && ci[1] < g.n[1] && (N == 2 || (lz < box.e[2] && ci[2] < g.n[2]));
        c = ok ? jp_cell_lin<N>(g, ci) : 0;
    } else ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    double a[3] = {0.0, 0.0, 0.0};                      // lower vertices of the cell
    if (ok) {
#pragma unroll
        for (int d = 0; d < N; d++) a[d] = g.xv[d][ci[d]];
    }
    uint64_t lv = 0, codew = 0;
    int k = 0;
    unsigned cplx = 0;     // reason bits: 1 far / on a vertex, 2 same cell (ulp gap), 4 fails isincell in destination
#ifndef JP_CLS_U
#define JP_CLS_U 4
#endif
    constexpr int U = JP_CLS_U;
    for (int s0 = 0; s0 < g.S; s0 += U) {
        const unsigned bits = (unsigned)(m >> s0) & ((1u << U) - 1u);
        if (!__any_sync(0xffffffffu, bits != 0)) continue;
        double p[U][3];
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int d = 0; d < N; d++) p[u][d] = ((bits >> u) & 1u) ? co.p[d][c + (int64_t)(s0 + u) * g.C] : 0.0;
        // range grids: single-precision pre-filter (jp_classify_fast), branch-free; the exact comparisons
        // (isincell with upper edge fl(a + dx), domain test, destination among the four vertices) run only
        // for batches holding a particle within 1e-4 dx of a vertex -- and always on vector grids
        int codes[U];
        bool unsure = false;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool live = (bits >> u) & 1u;
            const int cf = g.cls_fast ? jp_classify_fast<N>(g, ci, a, p[u]) : -1;
            codes[u] = live ? cf : JP_CLS_STAY;
            unsure = unsure || codes[u] < 0;
        }
        if (__any_sync(0xffffffffu, unsure)) {
            if (unsure) {
                // the four vertices around the cell per dimension (NaN outside the grid: comparisons fail)
                double am[3], b[3], bp[3];
#pragma unroll
                for (int d = 0; d < N; d++) {
                    const double *xv = g.xv[d];
                    const int i = ci[d];
                    b[d] = xv[i + 1];
                    am[d] = i > 0 ? xv[i - 1] : NAN;
                    bp[d] = i + 2 <= g.n[d] ? xv[i + 2] : NAN;
                }
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (codes[u] < 0) codes[u] = jp_classify_particle<N>(g, am, a, b, bp, p[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            int code = codes[u];
            if (code == JP_CLS_STAY) continue;
            lv |= 1ull << (s0 + u);
            if (code > JP_CLS_CPLX) { cplx |= 1u << (code - JP_CLS_CPLX - 1); code = JP_CODE_DELETE; }
            codew |= (uint64_t)code << (8 * (k & 7));
            if ((++k & 7) == 0) { ws.code[(int64_t)((k >> 3) - 1) * g.C + c] = codew; codew = 0; }
        }
    }
    if (ok) {
        ws.occ[c] = m; ws.occ0[c] = m; ws.leave[c] = lv;
        if (k & 7) ws.code[(int64_t)(k >> 3) * g.C + c] = codew;
    }
    const unsigned wc = __reduce_or_sync(0xffffffffu, cplx);
    if (wc && threadIdx.x == 0) atomicOr(complex_flag, wc);
}

// ---- JP_MOVE_POLICY_DENSE ("vacate everything, then place"; opt-in, not reference behaviour): every cell gives up its leavers'
// slots before the first migrant is placed, so a migrant finds the slots its destination vacates in the same call.
__global__ void __launch_bounds__(256) k_move_prevacate(int64_t C, MovePlanWs ws, const unsigned int *__restrict__ skip_flag) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C || *skip_flag) return;
    ws.occ[c] = ws.occ0[c] & ~ws.leave[c];
}

// (Tried: prefetch.global.L1 of the neighbourhood's occupancy rows before the dependent chain starts -- plan
// 1.41 -> 1.72 ms at 256^3, 0.38 -> 0.43 ms at 128^3 (profiles/r02ag_ab_plan_prefetch.log) -- 18 more instructions per source cell,
// wasted on the cells without leavers and the rows nobody asks for.)
// ---- B. one colour of the plan (thread = source cell; 8-byte words only).
// Literal slot logic of move_kernel! (src/Particles/move_safe.jl:72-125) on the occupancy
// words; the slot given to the k-th leaver goes to res (one byte: slot | placed << 6).
template <int N>
__device__ __forceinline__ void jp_move_plan_cell(const JpGrid &g, const MovePlanWs &ws, int ox, int oy, int oz, int ncx, int ncy, int64_t t,
                                                  long long *stats, int policy) {
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
    const bool compact = policy != JP_MOVE_POLICY_REFERENCE, dense = policy == JP_MOVE_POLICY_DENSE;
    while (lv) {
        const int ip = __ffsll((long long)lv) - 1;
        lv &= lv - 1;
        if ((k & 7) == 0) {
            codew = ws.code[(int64_t)(k >> 3) * g.C + c];
            if (k > 0) { ws.res[(int64_t)((k >> 3) - 1) * g.C + c] = resw; resw = 0; }
        }
        const int code = (int)((codew >> (8 * (k & 7))) & 255);
        const int kk = k++;
        if (!dense) occ_c &= ~(1ull << ip);   // (DENSE: k_move_prevacate cleared it, and an earlier colour's migrant may sit there by now)
        if (code == JP_CODE_DELETE) { n_deleted++; continue; }
        int dv[3];
        jp_code_dir(code, dv);
        const int64_t c2 = c + dv[0] + (int64_t)g.n[0] * (dv[1] + (N == 3 ? (int64_t)g.n[1] * dv[2] : 0));
        const uint64_t o2 = ws.occ[c2];
        const uint64_t freebits = ~o2 & smask & (~0ull << cursor);
        if (freebits == 0) { n_dropped++; continue; }
        const int fs = __ffsll((long long)freebits) - 1;
        if (!compact) cursor = fs;            // JP_MOVE_POLICY_COMPACT / _DENSE: every search starts at slot 0
        ws.occ[c2] = o2 | (1ull << fs);
        resw |= (uint64_t)(fs | 64) << (8 * (kk & 7));
    }
    if (!dense) ws.occ[c] = occ_c;
    ws.res[(int64_t)((k - 1) >> 3) * g.C + c] = resw;
    if (n_dropped) atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_dropped);
    if (n_deleted) atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_deleted);
}

// One colour per launch: grids whose colours are larger than what is co-resident (a grid-stride loop would serialise the dependent
// loads of two cells per thread: 1.43 -> 1.93 ms at 256^3).
template <int N>
__global__ void __launch_bounds__(256) k_move_plan(JpGrid g, MovePlanWs ws, int ox, int oy, int oz, int ncx, int ncy, int64_t ncol,
                                                   long long *stats, int policy, const unsigned int *__restrict__ skip_flag) {
    if (*skip_flag) return;                                   // the call takes the direct sweeps (decided on the device)
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ncol) jp_move_plan_cell<N>(g, ws, ox, oy, oz, ncx, ncy, t, stats, policy);
}

// All 3^N colours in ONE cooperative launch (grid-wide barrier between colours, grid-stride within a colour) instead of 3^N launches:
// on a small grid a colour is a few microseconds of work behind a launch, and on any grid the 3^N - 1 launch gaps go away.
// The flag is not written while this kernel runs, so the early return is taken by every thread or by none.
template <int N>
__global__ void __launch_bounds__(256) k_move_plan_all(JpGrid g, MovePlanWs ws, int ncx, int ncy, int64_t ncol,
                                                       long long *stats, int policy, const unsigned int *__restrict__ skip_flag) {
    if (*skip_flag) return;                                   // the call takes the direct sweeps (decided on the device)
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool first = true;
    for (int ox = 0; ox < 3; ox++)
        for (int oy = 0; oy < 3; oy++)
            for (int oz = 0; oz < (N == 3 ? 3 : 1); oz++) {
                if (!first) grid.sync();
                first = false;
                for (int64_t t = tid; t < ncol; t += nthr) jp_move_plan_cell<N>(g, ws, ox, oy, oz, ncx, ncy, t, stats, policy);
            }
}

// ---- C. arrival mask + count per cell: every slot occupied at the end that is not a
// non-leaving original occupant holds an arrival.
template <int N>
__global__ void __launch_bounds__(256) k_move_finalize(JpGrid g, MovePlanWs ws, const unsigned int *__restrict__ skip_flag) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.C) return;
    if (*skip_flag) { ws.cnt[c] = 0; return; }
    const uint64_t am = ws.occ[c] & (~ws.occ0[c] | ws.leave[c]);
    ws.arrmask[c] = am;
    ws.cnt[c] = (uint32_t)__popcll(am);
}

struct MoveArrays { double *a[JP_MAX_ARGS + 3]; int n; };

// ---- D. gather (source-centric, slot-synchronous => every source sector is read once, in
// streaming order): payload of each placed leaver -> staging[off[dest] + rank of its slot].
// Slots are handled in batches of U with all loads of a batch issued before the stores.
#ifndef JP_MV_U
#define JP_MV_U 4
#endif
// (Tried on gather and fused scatter: single-use data -- leaver payload, staging records, stayers read for the interpolation sums --
// marked evict-first (ld.global.cs / st.global.cs) so that the partially written sectors stay in the L2 longer: gather 6.86 -> 7.44 ms,
// fused scatter 13.8 -> 17.2 ms at 256^3, profiles/r02af_ab_stream_hints.log; the .cs loads lose the L1 lines the prefetches bring in.
// prefetch.global.L1 of the leavers' sectors a batch ahead in the gather: +0.5 ms, profiles/r02n_ab_gather_prefetch.log.)
#define JP_MV_A 4      // arrays handled per register batch (coords + fields); more arrays loop again
template <int N>
__global__ void __launch_bounds__(256, JP_MINB_GATHER) k_move_gather(JpGrid g, MovePlanWs ws, MoveArrays arrs, double *__restrict__ stage,
                                                                   const unsigned int *__restrict__ skip_flag) {
    if (*skip_flag) return;
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t lv = ok ? ws.leave[c] : 0;
    if (!__any_sync(0xffffffffu, lv != 0)) return;
    uint64_t codew = 0, resw = 0;
    int respl = -1;
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
                if ((k >> 3) != respl) { respl = k >> 3; resw = ws.res[(int64_t)respl * g.C + c]; codew = ws.code[(int64_t)respl * g.C + c]; }
                const int r = (int)((resw >> (8 * (k & 7))) & 255);
                if (r & 64) {
                    const int fs = r & 63;
                    const int code = (int)((codew >> (8 * (k & 7))) & 255);
                    int dv[3];
                    jp_code_dir(code, dv);
                    const int64_t c2 = c + dv[0] + (int64_t)g.n[0] * (dv[1] + (N == 3 ? (int64_t)g.n[1] * dv[2] : 0));
                    pos[u] = (int64_t)ws.off[c2] + __popcll(ws.arrmask[c2] & ((1ull << fs) - 1));
                    act[u] = true;
                }
            }
        }
        // staging is array-of-structs: one migrant = AS consecutive doubles (AS = n rounded up to 4,
        // padding written too), so every touched 32-byte sector is written completely (no fill read)
        const int AS = (arrs.n + 3) & ~3;
        for (int a0 = 0; a0 < AS; a0 += JP_MV_A) {
            double v[JP_MV_U][JP_MV_A];
#pragma unroll
            for (int u = 0; u < JP_MV_U; u++)
#pragma unroll
                for (int a = 0; a < JP_MV_A; a++)
                    v[u][a] = (act[u] && a0 + a < arrs.n) ? arrs.a[a0 + a][e[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < JP_MV_U; u++)
                if (act[u]) {
                    double2 *dst = reinterpret_cast<double2 *>(stage + pos[u] * AS + a0);
                    dst[0] = make_double2(v[u][0], v[u][1]);
                    dst[1] = make_double2(v[u][2], v[u][3]);
                }
        }
    }
}

// ---- E. scatter: arrivals from staging, NaN into vacated slots, mask bytes.
// (Measured: walking each thread's own arrivals in rank order instead -- more loads in flight, but every
// 8-byte store then goes to the L2 alone instead of merged with its x-neighbours' stores to the same
// 32-byte sector -- is 3x SLOWER; the slot-synchronous order stays.  prefetch.global.L2 of the cell's
// staging records ahead of the sweep: +4 % time; of the leavers' sectors in the gather: +80 %.)
template <int N>
__global__ void __launch_bounds__(256, JP_MINB_SCATTER) k_move_scatter(JpGrid g, MovePlanWs ws, MoveArrays arrs, uint8_t *index, const double *__restrict__ stage,
                                                                     const unsigned int *__restrict__ skip_flag) {
    if (*skip_flag) return;
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
        const int AS = (arrs.n + 3) & ~3;
        for (int a0 = 0; a0 < arrs.n; a0 += JP_MV_A) {
            double v[JP_MV_U][JP_MV_A];
#pragma unroll
            for (int u = 0; u < JP_MV_U; u++)
#pragma unroll
                for (int a = 0; a < JP_MV_A; a++)
                    if (a0 + a < arrs.n && ((chb >> u) & 1u)) v[u][a] = ((arb >> u) & 1u) ? stage[pos[u] * AS + a0 + a] : NAN;
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

// after the scan: total number of arrivals -> stats / read-back word; a staging buffer that is too small sends the call
// to the direct sweeps (flag bit 8) -- the host learns the count asynchronously and grows the buffer for the next call
#define JP_CPLX_STAGING 8u
__global__ void k_move_after_scan(long long *stats, const uint32_t *total, uint64_t cap_rows, unsigned int *flag, unsigned int *m_out) {
    const unsigned int f = *flag;
    if (f) return;
    const uint32_t M = *total;
    *m_out = M;
    if ((uint64_t)M > cap_rows) {
        *flag = JP_CPLX_STAGING;
        stats[0] = 0; stats[1] = 0; stats[2] = 0;       // the plan kernels counted drops / deletions: the direct sweeps count again
        return;
    }
    stats[0] = (long long)M;
}
