// justpic_sm100a.cu -- CUDA kernels (sm_100a) + C ABI of libjustpic_sm100a.so.
// See include/justpic_c.h for the boundary and DESIGN.md for the kernel notes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        (no implicit FMA contraction: fma() is explicit where the reference fuses).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <cub/device/device_scan.cuh>

#include "../../../include/justpic_c.h"
#include "jp_core.h"
#include "jp_host_grid.h"

// ---------------------------------------------------------------------------
// error plumbing
static thread_local char g_err[512] = "";
static int jp_fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define JP_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) return jp_fail(JP_ERR_CUDA, #call ": %s", cudaGetErrorString(e__)); \
    } while (0)
#define JP_CHECK_LAUNCH()                                                              \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) return jp_fail(JP_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e__)); \
    } while (0)

extern "C" const char *jp_last_error(void) { return g_err; }
extern "C" int jp_version(void) { return 100; }

struct MovePlanWs {
    uint64_t *occ;        // [C] running occupancy (plan) -> final occupancy
    uint64_t *occ0;       // [C] occupancy before the move
    uint64_t *leave;      // [C] slots vacated by the move (original leavers)
    uint64_t *code;       // [ceil(S/8)][C] destination codes of the leavers, one byte each, slot order
    uint64_t *res;        // [ceil(S/8)][C] results (dest slot | placed << 6), one byte per leaver, slot order
    uint64_t *arrmask;    // [C] slots receiving an arrival
    uint32_t *cnt;        // [C+1] arrivals per cell
    uint32_t *off;        // [C+1] exclusive scan of cnt (off[C] = total)
};

struct jp_ctx {
    int device;
    JpGrid g;             // device pointers
    void *gridmem;        // one allocation holding all grid vectors
    uint64_t *occ, *leave;  // [C] occupancy / leave words (move, inject)
    uint8_t *inbox;       // [C] 1: every live particle of the cell lies in its closed box (inject)
    int *inj_list;        // [2^N][cap] cells to inject into, per colour
    unsigned int *inj_count;  // [8]
    int64_t inj_cap;
    long long *stats;     // device counters: [0..2] move, [3] inject
    double *p2g_ws;       // [2 * 2^N * C] per-cell partial sums of the two-pass particle2grid (lazy)
    int p2g_mode;         // JP_P2G_EXACT / JP_P2G_TWOPASS / JP_P2G_TWOPASS_FASTW
    int move_policy;      // JP_MOVE_POLICY_REFERENCE (carried-over free-slot cursor) / JP_MOVE_POLICY_COMPACT / JP_MOVE_POLICY_DENSE
    int move_mode;        // JP_MOVE_AUTO (plan/gather/scatter, direct sweeps on ties) / JP_MOVE_DIRECT
    int affine_detected;  // grid vectors are exactly affine (jp_grid_build); g.affine = detected && option
    int last_move_path;   // 0 = plan, 1 = direct (diagnostics)
    int last_complex;     // reason bits of the last classification (diagnostics)
    MovePlanWs mp;        // plan workspace (lazy); occ/leave alias the fields above
    unsigned int *mp_flag;   // device: "complex" flag
    void *cub_tmp; size_t cub_tmp_bytes;
    double *stage; size_t stage_elems;   // staging buffer (grow-only)
    double *pr_ws; size_t pr_ws_elems;   // per-cell partial sums of the fused update_phase_ratios (lazy, grow-only)
    unsigned int *h_pinned;  // pinned host scratch for flag / totals
    // advection -> move hand-off (JP_OPT_ADVECT_CLASSIFY): the tiled advection kernel leaves the move plan's
    // classification words (what k_move_classify3 computes), valid only for the coordinate / index arrays
    // they were computed from and until the next library call that changes particles or uses the occupancy
    // words; planes rewritten by jp_halo_unpack are re-classified from the coordinates (hint_dirty)
    int hint_opt;            // option value (default 0)
    int hint_valid;          // mp.occ / occ0 / leave / code and mp_flag hold the classification of hint_key
    const void *hint_key[4]; // coords[0..2], index the words belong to
    int hint_ndirty; int hint_dirty[8][2];   // (dim, plane) rewritten since the hand-off
    int last_classify;       // 0: coordinates (k_move_classify3), 1: hand-off bytes (diagnostics)
    int adv_split;           // jp_advect_region: the shell part has run, the interior part is still to come
    int adv_split_bucketed;  // ... and the bucketed state that call saw (a halo unpack between the two parts touches shell cells only)
    int coop_plan_max, coop_sweep_max;   // co-resident blocks of the two cooperative move kernels (lazy)
    int capturing;           // the stream of the current call is being captured into a CUDA graph (set by PREP): no allocation, no event query, no host read
    int bucketed;            // every live particle lies strictly inside its storage cell (set by move / init / inject / clean, cleared by advect, halo unpack, foreign writes)
    int mp_ready;            // every buffer of the plan workspace is allocated
    void *last_stream;       // stream of the last jp_move (jp_last_move_path reads the device flag on it)
    cudaEvent_t m_event; int m_pending, m_probed;   // asynchronous read-back of the last arrival count (sizes the staging buffer)
    // move -> interpolation hand-off (JP_OPT_MOVE_INTERP, csrc/jp_move_interp.cuh): the scatter pass of jp_move also leaves the
    // two-pass particle2grid! partial sums of field mi_fp and the centre phase ratios of field mi_ph; valid for the arrays in
    // mi_key until the next library call that changes particles or those fields
    int mi_opt;
    const double *mi_fp, *mi_ph; int mi_K;          // registered by jp_move_interp_fields
    int mi_valid_p2g, mi_valid_ph, mi_p2g_mode;
    int last_p2g_handoff, last_phase_handoff;       // diagnostics: the last particle2grid! / phase_ratios_center! used the hand-off
    const void *mi_key[4];
    double *mi_rc; size_t mi_rc_elems;              // [K][C] centre ratios left by the scatter
    void *halo_buf; size_t halo_buf_bytes;          // jp_halo_exchange: 4 face buffers (send / receive, left / right), grow-only
    // JP_OPT_PROFILE: CUDA events around the stages of the planned jp_move (ring of the last JP_PROF_RING calls), read by jp_profile_read
    int prof_opt; cudaEvent_t *prof_ev; int prof_calls;
};
static inline void hint_invalidate(jp_ctx *ctx) { ctx->hint_valid = 0; ctx->hint_ndirty = 0; }
static void move_plan_free(jp_ctx *ctx);
#define JP_PROF_RING 32
#define JP_PROF_MARKS 6      // classify | plan | finalize + scan | gather | scatter | end
static inline void mi_invalidate(jp_ctx *ctx) { ctx->mi_valid_p2g = 0; ctx->mi_valid_ph = 0; }
// every entry point that changes particles (or may change particle fields) drops both hand-offs
static inline void handoffs_invalidate(jp_ctx *ctx) { hint_invalidate(ctx); mi_invalidate(ctx); }

struct Ptr3 { double *p[3]; };
struct CPtr3 { const double *p[3]; };

#include "jp_advect_tile.cuh"

// ---------------------------------------------------------------------------
// thread <-> cell mapping shared by all cell kernels: 32 x 8 tiles, z = blockIdx.z
#define JP_BX 32
#define JP_BY 8
// resident 256-thread CTAs per SM each streaming kernel sizes its registers for (measured on B200:
// classify / gather gain from occupancy, scatter and the two-pass particle2grid from registers)
#ifndef JP_MINB_CLASSIFY
#define JP_MINB_CLASSIFY 4
#endif
#ifndef JP_MINB_GATHER
#define JP_MINB_GATHER 4
#endif
#ifndef JP_MINB_SCATTER
#define JP_MINB_SCATTER 3
#endif
#ifndef JP_MINB_P2G
#define JP_MINB_P2G 2
#endif
template <int N>
__device__ __forceinline__ bool tile_cell(const JpGrid &g, int *ci, int64_t &c) {
    ci[0] = blockIdx.x * JP_BX + threadIdx.x;
    ci[1] = blockIdx.y * JP_BY + threadIdx.y;
    ci[2] = N == 3 ? blockIdx.z : 0;
    const bool ok = ci[0] < g.n[0] && ci[1] < g.n[1];
    c = ok ? jp_cell_lin<N>(g, ci) : 0;
    return ok;
}
static dim3 tile_grid(int nx, int ny, int nz) { return dim3((nx + JP_BX - 1) / JP_BX, (ny + JP_BY - 1) / JP_BY, nz); }

// live mask of a cell: bit s = index[c + s*C] != 0   (S independent byte loads in flight)
__device__ __forceinline__ uint64_t load_mask(const uint8_t *__restrict__ index, int64_t c, int64_t C, int S, bool ok) {
    uint64_t m = 0;
    if (ok) {
#pragma unroll 8
        for (int s = 0; s < S; s++) m |= (uint64_t)(index[c + (int64_t)s * C] != 0) << s;
    }
    return m;
}

#include "jp_move_plan.cuh"
#include "jp_convert.cuh"

// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(256) k_init(JpGrid g, Ptr3 co, uint8_t *index, int npq, uint64_t seed) {
    int ci[3]; int64_t c;
    if (!tile_cell<N>(g, ci, c)) return;
    jp_init_cell<N>(g, co.p, index, npq, seed, c);
}

// advection!: thread = cell, slot planes in lock-step across the warp (coalesced
// 256 B per coordinate load), planes that are dead for the whole warp skipped by ballot.
template <int N, int SCHEME, bool FAST, bool UNIFORM>
__global__ void __launch_bounds__(256) k_advect(JpGrid g, Ptr3 co, const uint8_t *__restrict__ index, CPtr3 V, double alpha, double dt) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    const int cell1[3] = {ci[0] + 1, ci[1] + 1, ci[2] + 1};
    for (int s = 0; s < g.S; s++) {
        const bool live = (m >> s) & 1ull;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (live) {
            const int64_t e = c + (int64_t)s * g.C;
            double p0[3], p1[3];
#pragma unroll
            for (int d = 0; d < N; d++) p0[d] = co.p[d][e];
            jp_advect_particle<N, SCHEME, FAST, UNIFORM>(g, alpha, V.p, dt, cell1, p0, p1);
#pragma unroll
            for (int d = 0; d < N; d++) co.p[d][e] = p1[d];
        }
    }
}

// advection_LinP! / advection_MQS!: same thread <-> cell sweep, literal LinP / MQS interpolants
template <int N, int SCHEME, int INTERP>
__global__ void __launch_bounds__(256) k_advect_hi(JpGrid g, Ptr3 co, const uint8_t *__restrict__ index, CPtr3 V, double alpha, double dt) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    const int cell1[3] = {ci[0] + 1, ci[1] + 1, ci[2] + 1};
    for (int s = 0; s < g.S; s++) {
        const bool live = (m >> s) & 1ull;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (live) {
            const int64_t e = c + (int64_t)s * g.C;
            double p0[3], p1[3];
#pragma unroll
            for (int d = 0; d < N; d++) p0[d] = co.p[d][e];
            jp_advect_particle_hi<N, SCHEME, INTERP>(g, alpha, V.p, dt, cell1, p0, p1);
#pragma unroll
            for (int d = 0; d < N; d++) co.p[d][e] = p1[d];
        }
    }
}

// move_particles! pass A: occupancy + leave words (order-free, coalesced)
template <int N>
__global__ void __launch_bounds__(256) k_move_classify(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, uint64_t *occ, uint64_t *leave,
                                                       const unsigned int *__restrict__ run_flag) {
    if (run_flag && !*run_flag) return;                       // JP_MOVE_AUTO: only when the planned path declined (device-side decision)
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    uint64_t lv = 0;
    double corner[3], dx[3];
    if (ok)
        for (int d = 0; d < N; d++) { corner[d] = g.xv[d][ci[d]]; dx[d] = jp_d_of(g.xv[d], g.uniform, ci[d]); }
    for (int s = 0; s < g.S; s++) {
        const bool live = (m >> s) & 1ull;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (live) {
            const int64_t e = c + (int64_t)s * g.C;
            double p[3];
#pragma unroll
            for (int d = 0; d < N; d++) p[d] = co.p[d][e];
            if (!jp_isincell<N>(p, corner, dx)) lv |= 1ull << s;
        }
    }
    if (ok) { occ[c] = m; leave[c] = lv; }
}

// move_particles! pass B: one colour of the 3^N sweeps.  One WARP per source
// cell: lanes = the cell's leaving particles (chunks of 32).  Coordinates, the
// destination cell (seeded bisection) and the destination occupancy words are
// fetched by all lanes in parallel; the order-dependent part of the reference
// loop (src/Particles/move_safe.jl:72-125: first free slot >= cursor, in slot
// order, cursor shared across destinations) runs as a warp-uniform loop over
// shuffled values and touches registers only; payloads then move in parallel.
// A cell holding a particle that fails isincell but bisects back into its own
// cell (exactly on a face / ulp gap) takes the serial literal routine.
__device__ __forceinline__ int nth_set_bit64(uint64_t m, int i) {
    const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    const int nlo = __popc(lo);
    return i < nlo ? (int)__fns(lo, 0, i + 1) : 32 + (int)__fns(hi, 0, i - nlo + 1);
}

template <int N>
__device__ __forceinline__ void jp_move_sweep_cell(const JpGrid &g, Ptr3 co, uint8_t *index, const JpArgs &args, uint64_t *occ, uint64_t *leave,
                                                   int ox, int oy, int oz, int ncx, int ncy, int64_t t, long long *stats, int compact) {
    const int lane = threadIdx.x & 31;
    int ci[3];
    ci[0] = 3 * (int)(t % ncx) + ox;
    ci[1] = 3 * (int)((t / ncx) % ncy) + oy;
    ci[2] = N == 3 ? 3 * (int)(t / ((int64_t)ncx * ncy)) + oz : 0;
    if (ci[0] >= g.n[0] || ci[1] >= g.n[1] || (N == 3 && ci[2] >= g.n[2])) return;
    const int64_t c = jp_cell_lin<N>(g, ci);
    uint64_t lv = leave[c];
    if (lv == 0) return;
    const int S = g.S;
    const uint64_t smask = S == 64 ? ~0ull : ((1ull << S) - 1);
    uint64_t occ_c = occ[c];
    int cursor = 0, n_moved = 0, n_dropped = 0, n_deleted = 0;
    double lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < N; d++) { lo[d] = g.xv[d][0]; hi[d] = g.xv[d][g.n[d]]; }
    while (lv) {
        const int n = min(__popcll(lv), 32);
        // ---- parallel: classify my particle
        const bool act = lane < n;
        int ip = 0;
        int64_t e = 0, c2 = -1;
        double p[3] = {0, 0, 0};
        bool indom = false, tie = false, fails_dest = false;
        uint64_t my_occ = 0;
        if (act) {
            ip = nth_set_bit64(lv, lane);
            e = c + (int64_t)ip * g.C;
            indom = true;
#pragma unroll
            for (int d = 0; d < N; d++) { p[d] = co.p[d][e]; indom = indom && (lo[d] < p[d] && p[d] < hi[d]); }
            if (indom) {
                int nc[3] = {0, 0, 0};
                double corner[3], dx[3];
#pragma unroll
                for (int d = 0; d < N; d++) {
                    nc[d] = jp_bisect1(p[d], g.xv[d], g.n[d] + 1, ci[d] + 1) - 1;
                    corner[d] = g.xv[d][nc[d]];
                    dx[d] = jp_d_of(g.xv[d], g.uniform, nc[d]);
                }
                c2 = jp_cell_lin<N>(g, nc);
                tie = c2 == c;
                fails_dest = !jp_isincell<N>(p, corner, dx);
                if (!tie) my_occ = occ[c2];
            }
        }
        if (__any_sync(0xffffffffu, tie)) {
            // rare: serial literal routine for the whole cell (nothing has been modified yet in this chunk)
            if (lane == 0) {
                leave[c] = lv; occ[c] = occ_c;
                int st[3] = {0, 0, 0};
                // the cursor of the chunks already done carries over: emulate by a cursor-aware call
                jp_move_cell<N>(g, co.p, index, args, occ, leave, c, ci, st, cursor, compact != 0);
                n_moved += st[0]; n_dropped += st[1]; n_deleted += st[2];
            }
            lv = 0;
            occ_c = 0;   // already stored by jp_move_cell
            __syncwarp();
            if (lane == 0) {
                if (n_moved) atomicAdd((unsigned long long *)&stats[0], (unsigned long long)n_moved);
                if (n_dropped) atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_dropped);
                if (n_deleted) atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_deleted);
            }
            return;
        }
        // ---- serial (warp-uniform): slot assignment in slot order
        int my_fs = -1;
        for (int k = 0; k < n; k++) {
            const int ipk = __shfl_sync(0xffffffffu, ip, k);
            const long long c2k = __shfl_sync(0xffffffffu, (long long)c2, k);
            const unsigned long long o2 = __shfl_sync(0xffffffffu, (unsigned long long)my_occ, k);
            occ_c &= ~(1ull << ipk);
            if (c2k < 0) { n_deleted++; continue; }
            const uint64_t freebits = ~o2 & smask & (~0ull << cursor);
            if (freebits == 0) { n_dropped++; continue; }
            const int fs = __ffsll((long long)freebits) - 1;
            if (!compact) cursor = fs;
            if (c2 == c2k) my_occ = o2 | (1ull << fs);
            if (lane == k) my_fs = fs;
            n_moved++;
        }
        // ---- parallel: move payloads
        if (act) {
            index[e] = 0;
#pragma unroll
            for (int d = 0; d < N; d++) co.p[d][e] = NAN;
            if (my_fs >= 0) {
                const int64_t e2 = c2 + (int64_t)my_fs * g.C;
                index[e2] = 1;
#pragma unroll
                for (int d = 0; d < N; d++) co.p[d][e2] = p[d];
                for (int a = 0; a < args.n; a++) { args.a[a][e2] = args.a[a][e]; args.a[a][e] = NAN; }
                occ[c2] = my_occ;                       // same final value from every lane targeting c2
                if (fails_dest) atomicOr((unsigned long long *)&leave[c2], 1ull << my_fs);
            } else {
                for (int a = 0; a < args.n; a++) args.a[a][e] = NAN;
            }
        }
        // consume the processed bits
        for (int k = 0; k < n; k++) lv &= lv - 1;
        __syncwarp();
    }
    if (lane == 0) {
        occ[c] = occ_c;
        leave[c] = 0;
        if (n_moved) atomicAdd((unsigned long long *)&stats[0], (unsigned long long)n_moved);
        if (n_dropped) atomicAdd((unsigned long long *)&stats[1], (unsigned long long)n_dropped);
        if (n_deleted) atomicAdd((unsigned long long *)&stats[2], (unsigned long long)n_deleted);
    }
}

// The launch: one warp per source cell, all 3^N colours in ONE cooperative launch (grid-wide barrier between colours, grid-stride
// within a colour).  In JP_MOVE_AUTO it sits behind the planned path and returns at once unless that declined (device-side flag;
// run_flag is not written while the kernel runs, so every thread takes the early return or none does).
template <int N>
__global__ void __launch_bounds__(256) k_move_sweep_all(JpGrid g, Ptr3 co, uint8_t *index, JpArgs args, uint64_t *occ, uint64_t *leave,
                                                        int ncx, int ncy, int64_t ncol, long long *stats, int compact,
                                                        const unsigned int *__restrict__ run_flag) {
    if (run_flag && !*run_flag) return;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    bool first = true;
    for (int ox = 0; ox < 3; ox++)
        for (int oy = 0; oy < 3; oy++)
            for (int oz = 0; oz < (N == 3 ? 3 : 1); oz++) {
                if (!first) grid.sync();
                first = false;
                for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < ncol; t += nwarps) {     // warp-uniform
                    jp_move_sweep_cell<N>(g, co, index, args, occ, leave, ox, oy, oz, ncx, ncy, t, stats, compact);
                    __syncwarp();
                }
            }
}

// force_injection! (src/Particles/forced_injection.jl:32-79): thread = cell; entry ip of p_new goes to slot ip if it is free
struct ForceVals { double v[JP_MAX_ARGS]; };
template <int N>
__global__ void __launch_bounds__(256) k_force_injection(JpGrid g, Ptr3 co, uint8_t *index, CPtr3 pnew, JpArgs fields, ForceVals vals) {
    int ci[3]; int64_t c;
    if (!tile_cell<N>(g, ci, c)) return;
    if (isnan(pnew.p[0][c])) return;                       // !isnan(p_new[I..., begin])
    for (int s = 0; s < g.S; s++) {
        const int64_t e = c + (int64_t)s * g.C;
        if (index[e]) continue;                            // doskip(index, ip, I...) || continue
#pragma unroll
        for (int d = 0; d < N; d++) co.p[d][e] = pnew.p[d][e];
        index[e] = 1;
        for (int a = 0; a < fields.n; a++) fields.a[a][e] = vals.v[a];
    }
}

template <int N>
__global__ void __launch_bounds__(256) k_clean(JpGrid g, Ptr3 co, uint8_t *index, JpArgs args) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    for (int s = 0; s < g.S; s++) {
        const bool live = (m >> s) & 1ull;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (live) {
            const int64_t e = c + (int64_t)s * g.C;
            double p[3];
#pragma unroll
            for (int d = 0; d < N; d++) p[d] = co.p[d][e];
            if (jp_clean_removes<N>(g, ci, p)) {
                index[e] = 0;
#pragma unroll
                for (int d = 0; d < N; d++) co.p[d][e] = NAN;
                for (int a = 0; a < args.n; a++) args.a[a][e] = NAN;
            }
        }
    }
}

// inject_particles! pass A (thread = cell, coalesced): occupancy word + "would the reference
// inject here?" -> the cell is appended to the work list of its colour (order within a colour
// is irrelevant: same-colour cells only read their neighbours, which have other colours).
template <int N, bool PHASE>
__global__ void __launch_bounds__(256) k_inject_classify(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, int min_xq,
                                                         uint64_t *occ, uint8_t *inbox, int *list, unsigned int *count, int64_t cap) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    double vq[3], ub[3], hi[3], dqv[3];
    if (ok)
        for (int d = 0; d < N; d++) { vq[d] = g.xv[d][ci[d]]; dqv[d] = jp_d_of(g.xv[d], g.uniform, ci[d]) / 2; ub[d] = vq[d] + dqv[d]; hi[d] = g.xv[d][ci[d] + 1]; }
    int nq0 = 0;
    uint64_t nq = 0;         // PHASE: live particles strictly inside each of the 2^N quadrants, one byte per quadrant
    bool allin = true;       // every live particle lies in the closed box [xv[i], xv[i+1]]^N (lets the donor search prune this cell)
    for (int s0 = 0; s0 < g.S; s0 += 4) {
        const unsigned bits = (unsigned)(m >> s0) & 15u;
        if (!__any_sync(0xffffffffu, bits != 0)) continue;
        double p[4][3];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int d = 0; d < N; d++) p[u][d] = ((bits >> u) & 1u) ? co.p[d][c + (int64_t)(s0 + u) * g.C] : NAN;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            bool in = (bits >> u) & 1u;
#pragma unroll
            for (int d = 0; d < N; d++) in = in && (vq[d] < p[u][d]) && (p[u][d] < ub[d]);
            nq0 += in ? 1 : 0;
            if (PHASE && ((bits >> u) & 1u)) {
                // quadrant (b_x, b_y[, b_z]): vertex vq = xv + dq*b, strict isincell(p, vq, dq) per dimension
                int q = 0;
                bool inq = true;
#pragma unroll
                for (int d = 0; d < N; d++) {
                    const double dq = dqv[d];
                    const double mid = vq[d] + dq * 1.0;
                    const bool lo = (vq[d] < p[u][d]) && (p[u][d] < vq[d] + dq), hi = (mid < p[u][d]) && (p[u][d] < mid + dq);
                    inq = inq && (lo || hi);
                    q |= (hi ? 1 : 0) << d;
                }
                if (inq) nq += 1ull << (8 * q);
            }
            if ((bits >> u) & 1u) {
#pragma unroll
                for (int d = 0; d < N; d++) allin = allin && (vq[d] <= p[u][d]) && (p[u][d] <= hi[d]);
            }
        }
    }
    if (ok) {
        occ[c] = m;
        inbox[c] = allin ? 1 : 0;
        bool cand = jp_inject_candidate(nq0, __popcll(m), g.S, min_xq);
        if (PHASE) {                                  // every quadrant is examined (injection.jl:271 `continue`)
            bool deficient = false;
#pragma unroll
            for (int q = 0; q < (N == 2 ? 4 : 8); q++) deficient = deficient || (int)((nq >> (8 * q)) & 255) < min_xq;
            cand = deficient && __popcll(m) < g.S;
        }
        if (cand) {
            const int col = (N == 3 ? ((ci[0] & 1) * 4 + (ci[1] & 1) * 2 + (ci[2] & 1)) : ((ci[0] & 1) * 2 + (ci[1] & 1)));
            const unsigned pos = atomicAdd(&count[col], 1u);
            list[(int64_t)col * cap + pos] = (int)c;
        }
    }
}

// inject_particles! pass B: one colour of the 2^N sweeps, ONE WARP per listed cell.
// Literal _inject_particles! (src/Particles/injection.jl:68-131): quadrant counts by ballot
// over the cell's slots, new particles in ascending free slots, and the nearest-donor search
// (index_min_distance, :330-393) spread over the lanes -- 3^N cells x S slots candidates,
// reduced with the lexicographic key (distance, visiting order) so that the winner is the
// candidate the reference's serial "strictly smaller" scan would keep.
// PHASE = inject_particles_phase! (src/Particles/injection.jl:146-325): every quadrant is examined, the
// nearest particle donates only its phase id, and args[j] is interpolated from the grid field
// fields[j] at the new position, clamped to the extrema of its interpolation stencil.
struct InjPhase { double *phases; const double *fields[JP_MAX_ARGS]; int32_t fkind[JP_MAX_ARGS]; };

template <int N>
__device__ __forceinline__ double jp_inject_field(const JpGrid &g, const double *__restrict__ F, int kind, const int *ci,
                                                  const double *xcc, const double *dcell, const double *pn) {
    int ic[3] = {0, 0, 0};
    double t[3], v[8];
    int64_t s1, s2;
    if (kind == 1) {                                   // centre field: shifted_index + clamp(., 1, n-1) (injection.jl:295-299)
#pragma unroll
        for (int d = 0; d < N; d++) {
            int i1 = ci[d] + 1;
            if (pn[d] < xcc[d]) i1 -= 1;
            const int hi1 = g.n[d] - 1;
            if (i1 > hi1) i1 = hi1; else if (i1 < 1) i1 = 1;
            ic[d] = i1 - 1;
            t[d] = (pn[d] - g.xc[d][ic[d]]) * (1.0 / jp_d_of(g.xc[d], g.uniform, ic[d]));
        }
        s1 = g.n[0]; s2 = (int64_t)g.n[0] * g.n[1];
    } else {                                           // vertex field of the storage cell (injection.jl:307-311)
#pragma unroll
        for (int d = 0; d < N; d++) { ic[d] = ci[d]; t[d] = (pn[d] - g.xv[d][ci[d]]) * (1.0 / dcell[d]); }
        s1 = g.n[0] + 1; s2 = (int64_t)(g.n[0] + 1) * (g.n[1] + 1);
    }
    jp_corners<N>(F, ic[0] + s1 * ic[1] + (N == 3 ? s2 * ic[2] : 0), s1, s2, v);
    const double tmp = jp_lerp<N>(v, t);
    double lo = v[0], hi = v[0];
#pragma unroll
    for (int q = 1; q < (N == 2 ? 4 : 8); q++) { lo = v[q] < lo ? v[q] : lo; hi = v[q] > hi ? v[q] : hi; }
    return tmp > hi ? hi : (tmp < lo ? lo : tmp);
}

// inject_particles! keys its random stream by (seed, step, cell, slot) with `step` an argument of the call.  A call captured into a
// CUDA graph would replay the same `step` for ever; so every inject kernel adds a device word (stats[7], 0 outside graphs), and a
// captured jp_inject / jp_inject_phase ends with k_step_advance: replay i of the graph injects with step + i, exactly what the
// eager loop that passes step, step + 1, ... does.  JP_OPT_GRAPH_STEP_OFFSET reads / resets the word.
__host__ __device__ inline unsigned int *jp_step_offset(long long *stats) { return reinterpret_cast<unsigned int *>(stats + 7); }
__global__ void k_step_advance(long long *stats) { *jp_step_offset(stats) += 1u; }

#ifndef JP_MINB_INJECT
#define JP_MINB_INJECT 2
#endif
template <int N, bool PHASE>
__device__ __forceinline__ void jp_inject_sweep_colour(const JpGrid &g, const Ptr3 &co, uint8_t *index, const JpArgs &args, uint64_t *occ,
                                                       uint8_t *inbox, const int *__restrict__ list, const unsigned n,
                                                       int min_xcell, uint64_t seed, uint32_t step, long long *stats, const InjPhase &ph) {
    const int lane = threadIdx.x & 31;
    const unsigned nwarps = gridDim.x * (blockDim.x >> 5);
    const int S = g.S, NQ = N == 2 ? 4 : 8;
    const int min_xq = (min_xcell + NQ - 1) / NQ;
    const uint64_t smask = S == 64 ? ~0ull : ((1ull << S) - 1);
    const int nx = g.n[0], ny = g.n[1], nz = N == 3 ? g.n[2] : 1;
    for (unsigned t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n; t += nwarps) {
        const int64_t c = list[t];
        int ci[3];
        jp_cell_ijk<N>(g, c, ci);
        uint64_t occ_c = occ[c];
        double p[2][3];
        bool live[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int s = lane + 32 * h;
            live[h] = s < S && ((occ_c >> s) & 1ull);
#pragma unroll
            for (int d = 0; d < N; d++) p[h][d] = live[h] ? co.p[d][c + (int64_t)s * g.C] : NAN;
        }
        double xvc[3], dq[3];
#pragma unroll
        for (int d = 0; d < N; d++) { xvc[d] = g.xv[d][ci[d]]; dq[d] = jp_d_of(g.xv[d], g.uniform, ci[d]) / 2; }
        // the 3^N neighbourhood, one neighbour per lane (fetched once per cell, broadcast by shuffle in the donor search below;
        // the neighbours have other colours, so their occupancy / pruning flag do not change during this launch)
        long long nb_c2 = -1;
        unsigned long long nb_occ = 0;
        int nb_inbox = 0;
        double nb_lo[3] = {0, 0, 0}, nb_hi[3] = {0, 0, 0};           // the neighbour's closed box (pruning)
        if (lane < (N == 3 ? 27 : 9)) {
            const int cc[3] = {ci[0] + lane % 3 - 1, ci[1] + (lane / 3) % 3 - 1, N == 3 ? ci[2] + lane / 9 - 1 : 0};
            if (cc[0] >= 0 && cc[1] >= 0 && cc[2] >= 0 && cc[0] < nx && cc[1] < ny && cc[2] < nz) {
                nb_c2 = cc[0] + (int64_t)nx * (cc[1] + (int64_t)ny * cc[2]);
                if (nb_c2 != c) { nb_occ = occ[nb_c2]; nb_inbox = inbox[nb_c2]; }
#pragma unroll
                for (int d = 0; d < N; d++) { nb_lo[d] = g.xv[d][cc[d]]; nb_hi[d] = g.xv[d][cc[d] + 1]; }
            }
        }
        int injected = 0;
#pragma unroll 1
        for (int iq = 0; iq < NQ; iq++) {
            double vq[3];
#pragma unroll
            for (int d = 0; d < N; d++) vq[d] = xvc[d] + dq[d] * (double)((iq >> d) & 1);
            int num = 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const bool in = live[h] && jp_isincell<N>(p[h], vq, dq);
                num += __popc(__ballot_sync(0xffffffffu, in));
            }
            if (num >= min_xq) { if (PHASE) continue; else break; }
            uint64_t freem = ~occ_c & smask;
            while (freem) {
                const int i = __ffsll((long long)freem) - 1;
                freem &= freem - 1;
                num++;
                double r[3], pn[3];
                jp_rand3(seed, PHASE ? 2u : 1u, step, (uint32_t)c, (uint32_t)i, r);
#pragma unroll
                for (int d = 0; d < N; d++) pn[d] = vq[d] + dq[d] * fma(0.95, r[d], 0.05);
                const int64_t e = c + (int64_t)i * g.C;
                if (lane == 0) {
                    bool inside = true;
#pragma unroll
                    for (int d = 0; d < N; d++) {
                        co.p[d][e] = pn[d];
                        inside = inside && (g.xv[d][ci[d]] <= pn[d]) && (pn[d] <= g.xv[d][ci[d] + 1]);
                    }
                    index[e] = 1;
                    if (!inside) inbox[c] = 0;       // keep the pruning flag truthful for later colours
                }
                occ_c |= 1ull << i;
                if ((i & 31) == lane) {
                    const int h = i >> 5;
                    if (h == 0) { live[0] = true; for (int d = 0; d < N; d++) p[0][d] = pn[d]; }
                    else        { live[1] = true; for (int d = 0; d < N; d++) p[1][d] = pn[d]; }
                }
                injected++;
                // nearest live particle in the 3^N neighbourhood.  The winner is the minimum of the
                // key (distance, visiting order of the reference), so cells may be evaluated in any
                // order: own cell first, then the others, skipping a cell when every particle it
                // holds is known to lie in its closed box (inbox) and the box is farther than the
                // best distance so far (all IEEE operations involved are monotonic, so the computed
                // box distance never exceeds the computed distance of a particle inside the box).
                // Every lane keeps its own best candidate (key: distance, then the reference's visiting order); ONE full
                // reduction at the end.  The pruning bound is the warp-wide minimum distance, refreshed (distance only) after each
                // visited cell.
                double best_d = INFINITY;
                int best_ord = 0x7fffffff;
                long long best_e = -1;
                const int self = N == 3 ? 13 : 4;
#pragma unroll
                for (int h = 0; h < 2; h++) {                       // own cell first, from registers
                    const int s = lane + 32 * h;
                    if (live[h] && s != i) {
                        const double dist = jp_distance<N>(p[h], pn);
                        const int ord = self * 64 + s;
                        if (dist < best_d || (dist == best_d && ord < best_ord)) { best_d = dist; best_ord = ord; best_e = c + (int64_t)s * g.C; }
                    }
                }
                double bound = best_d;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) { const double od = __shfl_xor_sync(0xffffffffu, bound, off); bound = od < bound ? od : bound; }
                // which neighbours can hold a closer particle: decided by all lanes at once (lane = neighbour): a cell whose particles
                // all lie in its closed box (inbox) is skipped when the box is farther than the bound (all IEEE operations involved are
                // monotonic, so the computed box distance never exceeds the computed distance of a particle inside the box)
                double bd = INFINITY;
                if (nb_c2 >= 0 && nb_c2 != c && nb_occ != 0) {
                    bd = 0.0;
                    if (nb_inbox) {
                        double bx[3];
#pragma unroll
                        for (int d = 0; d < N; d++) bx[d] = pn[d] < nb_lo[d] ? nb_lo[d] : (pn[d] > nb_hi[d] ? nb_hi[d] : pn[d]);      // nearest point of the box
                        bd = jp_distance<N>(bx, pn);
                    }
                }
                unsigned vis = __ballot_sync(0xffffffffu, !(bd > bound));
#pragma unroll 1
                while (vis) {
                    const int nidx = __ffs((int)vis) - 1;
                    vis &= vis - 1;
                    if (__shfl_sync(0xffffffffu, bd, nidx) > bound) continue;          // the bound has tightened since the ballot
                    const long long c2 = __shfl_sync(0xffffffffu, nb_c2, nidx);
                    const unsigned long long o2 = __shfl_sync(0xffffffffu, nb_occ, nidx);
                    double dmin = INFINITY;
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int s = lane + 32 * h;
                        if (s < S && ((o2 >> s) & 1ull)) {
                            double q[3];
#pragma unroll
                            for (int d = 0; d < N; d++) q[d] = co.p[d][c2 + (int64_t)s * g.C];
                            const double dist = jp_distance<N>(q, pn);
                            const int ord = nidx * 64 + s;
                            if (dist < best_d || (dist == best_d && ord < best_ord)) { best_d = dist; best_ord = ord; best_e = c2 + (int64_t)s * g.C; }
                            dmin = dist < dmin ? dist : dmin;
                        }
                    }
                    dmin = dmin < bound ? dmin : bound;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) { const double od = __shfl_xor_sync(0xffffffffu, dmin, off); dmin = od < dmin ? od : dmin; }
                    bound = dmin;
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, best_d, off);
                    const int oo = __shfl_xor_sync(0xffffffffu, best_ord, off);
                    const long long oe = __shfl_xor_sync(0xffffffffu, best_e, off);
                    if (od < best_d || (od == best_d && oo < best_ord)) { best_d = od; best_ord = oo; best_e = oe; }
                }
                if (PHASE) {
                    if (best_e >= 0 && lane == 0) ph.phases[e] = ph.phases[best_e];
                    if (lane < args.n) {                       // one lane per particle field
                        double xcc[3], dcell[3];
#pragma unroll
                        for (int d = 0; d < N; d++) { dcell[d] = jp_d_of(g.xv[d], g.uniform, ci[d]); xcc[d] = (xvc[d] + dq[d] * 0.0) + dq[d]; }
                        args.a[lane][e] = jp_inject_field<N>(g, ph.fields[lane], ph.fkind[lane], ci, xcc, dcell, pn);
                    }
                } else if (best_e >= 0 && lane == 0)
                    for (int a = 0; a < args.n; a++) args.a[a][e] = args.a[a][best_e];
                __syncwarp();
                if (num >= min_xq) break;
            }
        }
        if (lane == 0) {
            occ[c] = occ_c;
            if (injected) atomicAdd((unsigned long long *)&stats[3], (unsigned long long)injected);
        }
        __syncwarp();
    }
}

// One colour per launch (work lists built by k_inject_classify): a persistent grid of warps walks the list.  (All 2^N colours in one
// cooperative launch with grid-wide barriers was measured and dropped: 0.097 -> 0.101 ms at 2-D 256^2, 1.54 -> 1.58 ms at 128^3.)
template <int N, bool PHASE>
__global__ void __launch_bounds__(256, JP_MINB_INJECT) k_inject_sweep(JpGrid g, Ptr3 co, uint8_t *index, JpArgs args, uint64_t *occ,
                                                      uint8_t *inbox, const int *__restrict__ list, const unsigned int *__restrict__ count,
                                                      int min_xcell, uint64_t seed, uint32_t step, long long *stats, InjPhase ph) {
    step += *jp_step_offset(stats);                         // 0, or the number of replays of the CUDA graph this launch is part of
    jp_inject_sweep_colour<N, PHASE>(g, co, index, args, occ, inbox, list, *count, min_xcell, seed, step, stats, ph);
}

// ---- max_xcell > JP_MAX_SLOTS ("wide" cells, e.g. the reference's tests with max_xcell = 80 / 150): move_particles! and
// inject_particles!(_phase!) as literal per-cell kernels on the index bytes (thread = cell of the colour being swept).
// Same results as the occupancy-word kernels (the GPU parity tests cover both); not tuned -- the word kernels
// are the product path for every max_xcell <= 64.
template <int N>
__global__ void __launch_bounds__(128) k_move_sweep_wide(JpGrid g, Ptr3 co, uint8_t *index, JpArgs args, int ox, int oy, int oz, int ncx, int ncy,
                                                         int64_t ncol, long long *stats, int compact) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncol) return;
    int ci[3];
    ci[0] = 3 * (int)(t % ncx) + ox;
    ci[1] = 3 * (int)((t / ncx) % ncy) + oy;
    ci[2] = N == 3 ? 3 * (int)(t / ((int64_t)ncx * ncy)) + oz : 0;
    if (ci[0] >= g.n[0] || ci[1] >= g.n[1] || (N == 3 && ci[2] >= g.n[2])) return;
    int st[3] = {0, 0, 0};
    jp_move_cell_wide<N>(g, co.p, index, args, jp_cell_lin<N>(g, ci), ci, st, compact != 0);
    if (st[0]) atomicAdd((unsigned long long *)&stats[0], (unsigned long long)st[0]);
    if (st[1]) atomicAdd((unsigned long long *)&stats[1], (unsigned long long)st[1]);
    if (st[2]) atomicAdd((unsigned long long *)&stats[2], (unsigned long long)st[2]);
}

// literal _inject_particles_phase! (src/Particles/injection.jl:205-325) for one cell, serial
template <int N>
__device__ int jp_inject_phase_cell(const JpGrid &g, double *const *coords, uint8_t *index, const JpArgs &args, const InjPhase &ph,
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
        if (num >= min_xq) continue;                      // every quadrant is examined (injection.jl:271)
        for (int i = 0; i < S; i++) {
            const int64_t e = c + (int64_t)i * C;
            if (index[e]) continue;
            num++;
            double r[3], pn[3];
            jp_rand3(seed, 2u, step, (uint32_t)c, (uint32_t)i, r);
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
            if (emin >= 0) ph.phases[e] = ph.phases[emin];
            double xcc[3], dcell[3];
            for (int d = 0; d < N; d++) { dcell[d] = jp_d_of(g.xv[d], g.uniform, ci[d]); xcc[d] = (xvc[d] + dq[d] * 0.0) + dq[d]; }
            for (int a = 0; a < args.n; a++) args.a[a][e] = jp_inject_field<N>(g, ph.fields[a], ph.fkind[a], ci, xcc, dcell, pn);
            if (num >= min_xq) break;
        }
    }
    return injected;
}

template <int N, bool PHASE>
__global__ void __launch_bounds__(128) k_inject_wide(JpGrid g, Ptr3 co, uint8_t *index, JpArgs args, int ox, int oy, int oz, int ncx, int ncy, int64_t ncol,
                                                     int min_xcell, uint64_t seed, uint32_t step, long long *stats, InjPhase ph) {
    step += *jp_step_offset(stats);                         // 0, or the number of replays of the CUDA graph this launch is part of
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncol) return;
    int ci[3];
    ci[0] = 2 * (int)(t % ncx) + ox;
    ci[1] = 2 * (int)((t / ncx) % ncy) + oy;
    ci[2] = N == 3 ? 2 * (int)(t / ((int64_t)ncx * ncy)) + oz : 0;
    if (ci[0] >= g.n[0] || ci[1] >= g.n[1] || (N == 3 && ci[2] >= g.n[2])) return;
    const int64_t c = jp_cell_lin<N>(g, ci);
    const int inj = PHASE ? jp_inject_phase_cell<N>(g, co.p, index, args, ph, min_xcell, seed, step, c, ci)
                          : jp_inject_cell<N>(g, co.p, index, args, min_xcell, seed, step, c, ci);
    if (inj) atomicAdd((unsigned long long *)&stats[3], (unsigned long long)inj);
}

// grid2particle!
template <int N>
__global__ void __launch_bounds__(256) k_g2p(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, double *__restrict__ Fp, const double *__restrict__ F) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    double v[8], xcorner[3], idx[3];
    if (ok) {
        const int64_t s1 = g.n[0] + 1, s2 = (int64_t)(g.n[0] + 1) * (g.n[1] + 1);
        jp_corners<N>(F, ci[0] + s1 * ci[1] + (N == 3 ? s2 * ci[2] : 0), s1, s2, v);
        for (int d = 0; d < N; d++) { xcorner[d] = g.xv[d][ci[d]]; idx[d] = 1.0 / jp_d_of(g.xv[d], g.uniform, ci[d]); }
    }
    for (int s = 0; s < g.S; s++) {
        const bool live = (m >> s) & 1ull;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (live) {
            const int64_t e = c + (int64_t)s * g.C;
            double p[3];
#pragma unroll
            for (int d = 0; d < N; d++) p[d] = co.p[d][e];
            Fp[e] = jp_g2p<N>(v, xcorner, idx, p);
        }
    }
}

// grid2particle_flip! (src/Interpolations/grid_to_particle.jl:125-173): PIC/FLIP blend; spacing = grid_size(xvi)
template <int N>
__global__ void __launch_bounds__(256) k_g2p_flip(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, double *__restrict__ Fp,
                                                  const double *__restrict__ F, const double *__restrict__ F0, double alpha) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    double v[8], v0[8], xcorner[3], idx[3];
    if (ok) {
        const int64_t s1 = g.n[0] + 1, s2 = (int64_t)(g.n[0] + 1) * (g.n[1] + 1);
        const int64_t b = ci[0] + s1 * ci[1] + (N == 3 ? s2 * ci[2] : 0);
        jp_corners<N>(F, b, s1, s2, v);
        jp_corners<N>(F0, b, s1, s2, v0);
        for (int d = 0; d < N; d++) { xcorner[d] = g.xv[d][ci[d]]; idx[d] = g.inv_dmin_v[d]; }
    }
    for (int s = 0; s < g.S; s++) {
        const bool live = (m >> s) & 1ull;
        if (!__any_sync(0xffffffffu, live)) continue;
        if (live) {
            const int64_t e = c + (int64_t)s * g.C;
            double p[3];
#pragma unroll
            for (int d = 0; d < N; d++) p[d] = co.p[d][e];
            const double Fpic = jp_g2p<N>(v, xcorner, idx, p), F0pic = jp_g2p<N>(v0, xcorner, idx, p);
            const double Fflip = Fp[e] + (Fpic - F0pic);
            Fp[e] = fma(Fpic, alpha, Fflip * (1.0 - alpha));
        }
    }
}

// ---- subgrid_diffusion! (src/Physics/subgrid_diffusion.jl:55-143), fused into two particle passes:
// pass 1 = memcopy_cellarray_kernel! (all slots) + grid2particle!/centroid2particle!(pT, T_grid) +
//          subgrid_diffusion_kernel! (live slots);   [particle2grid!/particle2centroid! and the grid kernel run in between]
// pass 2 = grid2particle!/centroid2particle!(pdT, dT_subgrid) + update_particle_temperature_kernel! (all slots).
// Per-slot results are those of the reference's separate kernels (same operations on the same values);
// each slot is read and written once per pass instead of once per kernel.
template <int N, bool CENTROID>
__global__ void __launch_bounds__(256) k_subgrid_pass1(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, double *__restrict__ pT,
                                                       double *__restrict__ pT0, double *__restrict__ pdT, const double *__restrict__ dt0,
                                                       const double *__restrict__ Tg, double d, double dt) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    if (!ok) return;
    double v[8], xcorner[3], idx[3];
    if (!CENTROID) {
        const int64_t s1 = g.n[0] + 1, s2 = (int64_t)(g.n[0] + 1) * (g.n[1] + 1);
        jp_corners<N>(Tg, ci[0] + s1 * ci[1] + (N == 3 ? s2 * ci[2] : 0), s1, s2, v);
        for (int dd = 0; dd < N; dd++) { xcorner[dd] = g.xv[dd][ci[dd]]; idx[dd] = 1.0 / jp_d_of(g.xv[dd], g.uniform, ci[dd]); }
    }
    for (int s = 0; s < g.S; s++) {
        const int64_t e = c + (int64_t)s * g.C;
        const double told = pT[e];
        double t0 = told;                                            // memcopy: pT0 <- pT (every slot)
        double p[3];
        bool nan = false;
#pragma unroll
        for (int dd = 0; dd < N; dd++) { p[dd] = co.p[dd][e]; nan |= isnan(p[dd]); }
        const bool live = index[e] != 0;
        // grid2particle! uses the mask, centroid2particle! the NaN test (centroid_to_particle.jl:36)
        double tnew = told;
        if (CENTROID ? !nan : live) { tnew = CENTROID ? jp_c2p<N>(g, Tg, ci, p) : jp_g2p<N>(v, xcorner, idx, p); pT[e] = tnew; }
        if (live) {                                                  // subgrid_diffusion_kernel! (:119-133)
            const double q = dt0[e];
            const double den = isnan(q) ? q : (q > 1.0e-9 ? q : 1.0e-9);
            const double dTi = (tnew - t0) * (1 - exp(-d * dt / den));
            t0 = t0 + dTi;
            pdT[e] = dTi;
        }
        pT0[e] = t0;
    }
}

template <int N, bool CENTROID>
__global__ void __launch_bounds__(256) k_subgrid_pass2(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, double *__restrict__ pT,
                                                       const double *__restrict__ pT0, double *__restrict__ pdT, const double *__restrict__ dTs) {
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    if (!ok) return;
    double v[8], xcorner[3], idx[3];
    if (!CENTROID) {
        const int64_t s1 = g.n[0] + 1, s2 = (int64_t)(g.n[0] + 1) * (g.n[1] + 1);
        jp_corners<N>(dTs, ci[0] + s1 * ci[1] + (N == 3 ? s2 * ci[2] : 0), s1, s2, v);
        for (int dd = 0; dd < N; dd++) { xcorner[dd] = g.xv[dd][ci[dd]]; idx[dd] = 1.0 / jp_d_of(g.xv[dd], g.uniform, ci[dd]); }
    }
    for (int s = 0; s < g.S; s++) {
        const int64_t e = c + (int64_t)s * g.C;
        double p[3];
        bool nan = false;
#pragma unroll
        for (int dd = 0; dd < N; dd++) { p[dd] = co.p[dd][e]; nan |= isnan(p[dd]); }
        double dTi = pdT[e];
        if (CENTROID ? !nan : index[e] != 0) { dTi = CENTROID ? jp_c2p<N>(g, dTs, ci, p) : jp_g2p<N>(v, xcorner, idx, p); pdT[e] = dTi; }
        pT[e] = pT0[e] + dTi;                                        // update_particle_temperature_kernel! (every slot)
    }
}

// update_dT_subgrid_kernel! (:135-138): dTsubgrid[I] = dT[I + 1] - dTsubgrid[I]
template <int N>
__global__ void __launch_bounds__(256) k_update_dT_subgrid(double *__restrict__ dTs, const double *__restrict__ dT, int n0, int n1, int n2, int m0, int m1) {
    const int i = blockIdx.x * JP_BX + threadIdx.x, j = blockIdx.y * JP_BY + threadIdx.y, k = N == 3 ? blockIdx.z : 0;
    if (i >= n0 || j >= n1 || k >= n2) return;
    const int64_t a = i + (int64_t)n0 * (j + (int64_t)n1 * k);
    const int64_t b = (i + 1) + (int64_t)m0 * ((j + 1) + (N == 3 ? (int64_t)m1 * (k + 1) : 0));
    dTs[a] = dT[b] - dTs[a];
}

// centroid2particle!: liveness = !any(isnan, coords)  (the reference does not read the mask here)
template <int N>
__global__ void __launch_bounds__(256) k_c2p(JpGrid g, CPtr3 co, double *__restrict__ Fp, const double *__restrict__ Fc) {
    int ci[3]; int64_t c;
    if (!tile_cell<N>(g, ci, c)) return;
    for (int s = 0; s < g.S; s++) {
        const int64_t e = c + (int64_t)s * g.C;
        double p[3];
        bool nan = false;
#pragma unroll
        for (int d = 0; d < N; d++) { p[d] = co.p[d][e]; nan |= isnan(p[d]); }
        if (nan) continue;
        Fp[e] = jp_c2p<N>(g, Fc, ci, p);
    }
}

// particle2grid!: thread = node; fixed summation order (k, j, i, slot)
template <int N>
__global__ void __launch_bounds__(256) k_p2g(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, double *__restrict__ F, const double *__restrict__ Fp) {
    const int in = blockIdx.x * JP_BX + threadIdx.x, jn = blockIdx.y * JP_BY + threadIdx.y, kn = N == 3 ? blockIdx.z : 0;
    const int nx = g.n[0], ny = g.n[1], nz = N == 3 ? g.n[2] : 1;
    if (in > nx || jn > ny) return;
    const double xn[3] = {g.xv[0][in], g.xv[1][jn], N == 3 ? g.xv[2][kn] : 0.0};
    double w = 0.0, wF = 0.0;
    for (int ko = (N == 3 ? -1 : 0); ko <= 0; ko++) {
        const int kc = kn + ko;
        if (N == 3 && (kc < 0 || kc >= nz)) continue;
        for (int jo = -1; jo <= 0; jo++) {
            const int jc = jn + jo;
            if (jc < 0 || jc >= ny) continue;
            for (int io = -1; io <= 0; io++) {
                const int ic = in + io;
                if (ic < 0 || ic >= nx) continue;
                const int64_t c = ic + (int64_t)nx * (jc + (int64_t)ny * kc);
                for (int s = 0; s < g.S; s++) {
                    const int64_t e = c + (int64_t)s * g.C;
                    if (!index[e]) continue;
                    double p[3];
#pragma unroll
                    for (int d = 0; d < N; d++) p[d] = co.p[d][e];
                    const double wi = jp_p2g_weight<N>(xn, p);
                    w += wi;
                    wF = fma(wi, Fp[e], wF);
                }
            }
        }
    }
    const int64_t nd = in + (int64_t)(nx + 1) * (jn + (N == 3 ? (int64_t)(ny + 1) * kn : 0));
    F[nd] = N == 2 ? wF / w : wF * (1.0 / w);
}

// particle2grid!, two-pass deterministic variant (JP_P2G_TWOPASS).
// Pass 1 (cell-centric, streams every particle exactly once): for each cell the
// partial sums  sum_w[q], sum_wF[q]  of its particles w.r.t. each of its 2^N corner
// nodes q, accumulated in slot order with the reference's per-particle weight
// (sqrt -> square -> inv) and the reference's muladd.  Pass 2 (node-centric gather,
// no atomics): each node adds the partials of its <= 2^N adjacent cells in the
// reference's (k, j, i) order.  Same terms as the reference, fixed order, but the
// association is (cell sums) + ... instead of one running chain: results agree with
// the reference to a few ulp (stated tolerance 1e-12), not bit-for-bit.
// reciprocal to <= 2 ulp: MUFU seed + two Newton steps (no correctly-rounded fix-up, no
// special-case branch).  Only used by JP_P2G_TWOPASS_FASTW, whose contract is the 1e-12 tolerance.
__device__ __forceinline__ double jp_rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

template <int N, bool FASTW>
__global__ void __launch_bounds__(256, JP_MINB_P2G) k_p2g_cell(JpGrid g, CPtr3 co, const uint8_t *__restrict__ index, const double *__restrict__ Fp,
                                                  double *__restrict__ PW, double *__restrict__ PWF, const unsigned int *__restrict__ run_flag) {
    constexpr int NQ = N == 2 ? 4 : 8;
    constexpr int U = 4;                       // slots per batch: loads of a batch are issued together
    if (run_flag && !*run_flag) return;        // move -> interpolation hand-off: the partial sums are already there
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t m = load_mask(index, c, g.C, g.S, ok);
    double xn[3][2];
    if (ok)
        for (int d = 0; d < N; d++) { xn[d][0] = g.xv[d][ci[d]]; xn[d][1] = g.xv[d][ci[d] + 1]; }
    double aw[NQ], awf[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) { aw[q] = 0.0; awf[q] = 0.0; }
    for (int s0 = 0; s0 < g.S; s0 += U) {
        const unsigned bits = (unsigned)(m >> s0) & ((1u << U) - 1u);
        if (!__any_sync(0xffffffffu, bits != 0)) continue;
        double pp[U][3], ff[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t e = c + (int64_t)(s0 + u) * g.C;
            const bool live = (bits >> u) & 1u;
#pragma unroll
            for (int d = 0; d < N; d++) pp[u][d] = live ? co.p[d][e] : 0.0;
            ff[u] = live ? Fp[e] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if ((bits >> u) & 1u) {
                double d2[3][2];
#pragma unroll
                for (int d = 0; d < N; d++) {
                    const double a0 = xn[d][0] - pp[u][d], a1 = xn[d][1] - pp[u][d];
                    d2[d][0] = a0 * a0; d2[d][1] = a1 * a1;
                }
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    double ss = d2[0][q & 1] + d2[1][(q >> 1) & 1];
                    if (N == 3) ss = ss + d2[2][(q >> 2) & 1];
                    double wi;
                    if (FASTW) wi = jp_rcp_fast(ss);
                    else { const double dist = sqrt(ss); wi = 1.0 / (dist * dist); }
                    aw[q] += wi;
                    awf[q] = fma(wi, ff[u], awf[q]);
                }
            }
        }
    }
    if (ok) {
#pragma unroll
        for (int q = 0; q < NQ; q++) { PW[(int64_t)q * g.C + c] = aw[q]; PWF[(int64_t)q * g.C + c] = awf[q]; }
    }
}

template <int N>
__global__ void __launch_bounds__(256) k_p2g_node(JpGrid g, const double *__restrict__ PW, const double *__restrict__ PWF, double *__restrict__ F) {
    const int in = blockIdx.x * JP_BX + threadIdx.x, jn = blockIdx.y * JP_BY + threadIdx.y, kn = N == 3 ? blockIdx.z : 0;
    const int nx = g.n[0], ny = g.n[1], nz = N == 3 ? g.n[2] : 1;
    if (in > nx || jn > ny) return;
    double w = 0.0, wF = 0.0;
    for (int ko = (N == 3 ? -1 : 0); ko <= 0; ko++) {
        const int kc = kn + ko;
        if (N == 3 && (kc < 0 || kc >= nz)) continue;
        for (int jo = -1; jo <= 0; jo++) {
            const int jc = jn + jo;
            if (jc < 0 || jc >= ny) continue;
            for (int io = -1; io <= 0; io++) {
                const int ic = in + io;
                if (ic < 0 || ic >= nx) continue;
                const int64_t c = ic + (int64_t)nx * (jc + (int64_t)ny * kc);
                const int q = (io < 0 ? 1 : 0) | (jo < 0 ? 2 : 0) | ((N == 3 && ko < 0) ? 4 : 0);
                w = w + PW[(int64_t)q * g.C + c];
                wF = wF + PWF[(int64_t)q * g.C + c];
            }
        }
    }
    const int64_t nd = in + (int64_t)(nx + 1) * (jn + (N == 3 ? (int64_t)(ny + 1) * kn : 0));
    F[nd] = N == 2 ? wF / w : wF * (1.0 / w);
}

// particle2centroid!
template <int N>
__global__ void __launch_bounds__(256) k_p2c(JpGrid g, CPtr3 co, double *__restrict__ Fc, const double *__restrict__ Fp) {
    int ci[3]; int64_t c;
    if (!tile_cell<N>(g, ci, c)) return;
    double xcn[3], idi[3];
    for (int d = 0; d < N; d++) { xcn[d] = g.xc[d][ci[d]]; idi[d] = 1.0 / jp_d_of(g.xv[d], g.uniform, ci[d]); }
    double w = 0.0, wF = 0.0;
    for (int s = 0; s < g.S; s++) {
        const int64_t e = c + (int64_t)s * g.C;
        double p[3];
#pragma unroll
        for (int d = 0; d < N; d++) p[d] = co.p[d][e];
        if (N == 2 ? (isnan(p[0]) || isnan(p[1])) : isnan(p[0])) continue;
        const double wi = jp_bilinear_weight<N>(xcn, p, idi);
        w += wi;
        wF = fma(wi, Fp[e], wF);
    }
    Fc[c] = N == 2 ? wF / w : wF * (1.0 / w);
}

// phase_ratios_center!  Liveness is the reference's isnan(px) test, so px of EVERY slot
// is read; slots are processed in batches of U with the loads of a batch in flight together.
template <int N, int KMAX>
__global__ void __launch_bounds__(256, KMAX <= 8 ? 3 : 1) k_phase(JpGrid g, CPtr3 co, double *__restrict__ ratios, const double *__restrict__ phases, int K,
                                                                  const unsigned int *__restrict__ run_flag) {
    constexpr int U = 4;
    if (run_flag && !*run_flag) return;        // move -> interpolation hand-off: the ratios are copied from the workspace instead
    int ci[3]; int64_t c;
    if (!tile_cell<N>(g, ci, c)) return;
    double xcn[3], idi[3], w[KMAX];
    for (int d = 0; d < N; d++) { xcn[d] = g.xc[d][ci[d]]; idi[d] = 1.0 / jp_d_of(g.xv[d], g.uniform, ci[d]); }
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = 0.0;
    for (int s0 = 0; s0 < g.S; s0 += U) {
        double px[U], py[U], pz[U], ph[U];
#pragma unroll
        for (int u = 0; u < U; u++) px[u] = (s0 + u < g.S) ? co.p[0][c + (int64_t)(s0 + u) * g.C] : NAN;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool live = !isnan(px[u]);
            const int64_t e = c + (int64_t)(s0 + u) * g.C;
            py[u] = live ? co.p[1][e] : 0.0;
            pz[u] = (N == 3 && live) ? co.p[2][e] : 0.0;
            ph[u] = live ? phases[e] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (isnan(px[u])) continue;
            const double p[3] = {px[u], py[u], pz[u]};
            const double x = jp_bilinear_weight<N>(xcn, p, idi);
#pragma unroll
            for (int k = 0; k < KMAX; k++)
                if (k < K) w[k] = w[k] + (ph[u] == (double)(k + 1) ? x : copysign(0.0, x));
        }
    }
    double sum = w[0];
#pragma unroll
    for (int k = 1; k < KMAX; k++) if (k < K) sum = sum + w[k];
    const double inv = 1.0 / sum;
#pragma unroll
    for (int k = 0; k < KMAX; k++) if (k < K) ratios[c + (int64_t)k * g.C] = w[k] * inv;
}

#include "jp_phase_ratios.cuh"
#include "jp_move_interp.cuh"

// move -> interpolation hand-off: centre ratios left by k_move_scatter_interp -> the caller's array (unless the move took the direct sweeps)
__global__ void __launch_bounds__(256) k_copy_unless(double *__restrict__ dst, const double *__restrict__ src, int64_t n, const unsigned int *__restrict__ skip_flag) {
    if (*skip_flag) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// update_cell_halo!: arrays of one exchange (kernels in jp_halo_nccl.cuh)
struct HaloArrs { double *a[JP_MAX_ARGS + 3]; int n; };

// ---------------------------------------------------------------------------
// host side
extern "C" int jp_ctx_create(const jp_grid_desc *d, int device, jp_ctx **out) {
    if (!d || !out) return jp_fail(JP_ERR_INVALID, "jp_ctx_create: null argument");
    JpGrid g;
    JpGridOffsets off;
    std::vector<double> h;
    const char *msg = jp_grid_build(d, g, h, off);
    if (msg) return jp_fail(strstr(msg, "max_xcell") ? JP_ERR_UNSUPPORTED : JP_ERR_INVALID, "jp_ctx_create: %s", msg);
    JP_CUDA(cudaSetDevice(device));
    jp_ctx *ctx = (jp_ctx *)calloc(1, sizeof(jp_ctx));
    ctx->device = device;
    double *dm = nullptr;
    cudaError_t e = cudaMalloc(&dm, h.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(dm, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->occ, g.C * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->leave, g.C * sizeof(uint64_t));
    ctx->inj_cap = (int64_t)((g.n[0] + 1) / 2) * ((g.n[1] + 1) / 2) * ((g.n[2] + 1) / 2);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->inj_list, sizeof(int) * 8 * ctx->inj_cap);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->inj_count, 8 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->inbox, g.C);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->stats, 8 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->stats, 0, 8 * sizeof(long long));
    if (e != cudaSuccess) {
        cudaFree(dm); cudaFree(ctx->occ); cudaFree(ctx->leave); cudaFree(ctx->inj_list); cudaFree(ctx->inj_count); cudaFree(ctx->inbox); cudaFree(ctx->stats);
        free(ctx);
        return jp_fail(JP_ERR_CUDA, "jp_ctx_create: %s", cudaGetErrorString(e));
    }
    ctx->gridmem = dm;
    jp_grid_rebase(g, off, dm);
    ctx->g = g;
    ctx->affine_detected = g.affine;
    ctx->p2g_mode = JP_P2G_TWOPASS_FASTW;
    *out = ctx;
    return JP_OK;
}

extern "C" void jp_ctx_destroy(jp_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->gridmem); cudaFree(ctx->occ); cudaFree(ctx->leave); cudaFree(ctx->inj_list); cudaFree(ctx->inj_count); cudaFree(ctx->inbox); cudaFree(ctx->stats);
    cudaFree(ctx->p2g_ws);
    move_plan_free(ctx);
    cudaFree(ctx->stage); cudaFree(ctx->pr_ws); cudaFree(ctx->mi_rc); cudaFree(ctx->halo_buf);
    if (ctx->prof_ev) { for (int i = 0; i < JP_PROF_RING * JP_PROF_MARKS; i++) cudaEventDestroy(ctx->prof_ev[i]); free(ctx->prof_ev); }
    free(ctx);
}

static int check_particles(const jp_ctx *ctx, const jp_particles *p, const char *who) {
    if (!ctx || !p) return jp_fail(JP_ERR_INVALID, "%s: null context/particles", who);
    for (int d = 0; d < ctx->g.ndim; d++)
        if (!p->coords[d]) return jp_fail(JP_ERR_INVALID, "%s: null coordinate array", who);
    if (!p->index) return jp_fail(JP_ERR_INVALID, "%s: null index array", who);
    return JP_OK;
}
static int pack_args(double *const *args, int nargs, JpArgs &out, const char *who) {
    if (nargs < 0 || nargs > JP_MAX_ARGS) return jp_fail(JP_ERR_UNSUPPORTED, "%s: 0 <= nargs <= 16 required", who);
    out.n = nargs;
    for (int a = 0; a < nargs; a++) {
        if (!args || !args[a]) return jp_fail(JP_ERR_INVALID, "%s: null particle field", who);
        out.a[a] = args[a];
    }
    for (int a = nargs; a < JP_MAX_ARGS; a++) out.a[a] = nullptr;
    return JP_OK;
}
// CUDA graphs: every entry point may be called on a stream that is being captured (cudaStreamBeginCapture / torch.cuda.graph).
// Nothing on the hot path reads the device back; what cannot be captured -- growing a workspace -- is refused with a message that
// says what to do (run the step once eagerly first), and jp_move skips its asynchronous staging-size feedback.
static inline int jp_stream_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return 0; }
    return cs == cudaStreamCaptureStatusActive;
}
#define JP_NO_ALLOC_IN_CAPTURE(what)                                                                                       \
    if (ctx->capturing)                                                                                                    \
        return jp_fail(JP_ERR_INVALID, "%s would have to be allocated while the stream is being captured into a CUDA graph: run the same calls once eagerly before capturing", what)

#define PREP(who)                                                    \
    int rc__ = check_particles(ctx, p, who);                         \
    if (rc__) return rc__;                                           \
    JP_CUDA(cudaSetDevice(ctx->device));                             \
    const JpGrid &g = ctx->g;                                        \
    cudaStream_t st = (cudaStream_t)stream;                          \
    ctx->capturing = jp_stream_capturing(st);                        \
    Ptr3 co = {{p->coords[0], p->coords[1], p->coords[2]}};          \
    CPtr3 cco = {{p->coords[0], p->coords[1], p->coords[2]}};        \
    const dim3 blk(JP_BX, JP_BY, 1);                                 \
    const dim3 grd = tile_grid(g.n[0], g.n[1], g.n[2]);              \
    (void)co; (void)cco; (void)blk; (void)grd; (void)st;

// max_xcell > JP_MAX_SLOTS: the per-slot kernels (advection, grid2particle, clean, ...) keep one 64-bit occupancy
// mask per cell, so they are launched once per chunk of 64 slot planes -- slots [64k, 64k + 64) of a CellArray
// are themselves a CellArray with the same cell stride, starting 64k*C elements further on.
struct SlotChunk { JpGrid g; int64_t off; };
static inline int jp_nchunks(const JpGrid &g) { return (g.S + JP_MAX_SLOTS - 1) / JP_MAX_SLOTS; }
static inline SlotChunk jp_chunk(const JpGrid &g, int ch) {
    SlotChunk k;
    k.g = g;
    k.g.S = g.S - JP_MAX_SLOTS * ch < JP_MAX_SLOTS ? g.S - JP_MAX_SLOTS * ch : JP_MAX_SLOTS;
    k.off = (int64_t)JP_MAX_SLOTS * ch * g.C;
    return k;
}
static inline Ptr3 jp_shift(Ptr3 a, int64_t off) { for (int d = 0; d < 3; d++) if (a.p[d]) a.p[d] += off; return a; }
static inline CPtr3 jp_shift(CPtr3 a, int64_t off) { for (int d = 0; d < 3; d++) if (a.p[d]) a.p[d] += off; return a; }
static inline JpArgs jp_shift(JpArgs a, int64_t off) { for (int i = 0; i < a.n; i++) a.a[i] += off; return a; }

extern "C" int jp_init_particles(jp_ctx *ctx, const jp_particles *p, int32_t nxcell, uint64_t seed, void *stream) {
    PREP("jp_init_particles");
    handoffs_invalidate(ctx);
    ctx->bucketed = 1;
    const int NQ = g.ndim == 2 ? 4 : 8;
    if (nxcell < 0) return jp_fail(JP_ERR_INVALID, "jp_init_particles: nxcell < 0");     // 0: empty container (test/test_2D.jl:302)
    const int npq = (nxcell + NQ - 1) / NQ;
    if (npq * NQ > g.S) return jp_fail(JP_ERR_INVALID, "jp_init_particles: nxcell (rounded up to a multiple of 2^N) exceeds max_xcell");
    if (g.ndim == 2) k_init<2><<<grd, blk, 0, st>>>(g, co, p->index, npq, seed);
    else             k_init<3><<<grd, blk, 0, st>>>(g, co, p->index, npq, seed);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver-entry-point lookup (no link against libcuda)
typedef CUresult (*jp_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static jp_encode_tiled_fn jp_get_encode_tiled() {
    static jp_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (getenv("JP_NO_TMA") == nullptr &&
            cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (jp_encode_tiled_fn)p;
    }
    return fn;
}

// Tensor maps for the velocity components that satisfy the TMA constraints; returns the component mask.
template <int N, int H = 0>
static int build_advect_tma(const JpGrid &g, CPtr3 V, AdvTmaMaps &maps) {
    using T = AdvTile<N, H>;
    memset(&maps, 0, sizeof(maps));
    jp_encode_tiled_fn enc = jp_get_encode_tiled();
    if (!enc) return 0;
    int mask = 0;
    for (int c = 0; c < N; c++) {
        const cuuint64_t dims[3] = {(cuuint64_t)g.nvel[c][0], (cuuint64_t)g.nvel[c][1], (cuuint64_t)(N == 3 ? g.nvel[c][2] : 1)};
        const cuuint64_t strides[2] = {dims[0] * 8, dims[0] * dims[1] * 8};        // bytes, dims 1..rank-1
        if (((uintptr_t)V.p[c] & 15) || (strides[0] & 15) || (N == 3 && (strides[1] & 15))) continue;
        const cuuint32_t box[3] = {(cuuint32_t)T::EX, (cuuint32_t)T::EY, (cuuint32_t)T::EZ};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&maps.m[c], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, N, (void *)V.p[c], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            mask |= 1 << c;
    }
    return mask;
}

static int move_plan_alloc(jp_ctx *ctx);
struct AdvHandoff { MovePlanWs ws; unsigned int *flag; };        // flag == nullptr: no hand-off
template <int N, int SCHEME, bool UNIFORM, int AFFINE, bool HINT, int INTERP, bool BKT>
static cudaError_t launch_advect_tile_k(const JpGrid &g, cudaStream_t st, Ptr3 co, const uint8_t *index, CPtr3 V, double alpha, double dt, const AdvHandoff &ho) {
    constexpr int H = INTERP ? 1 : 0;
    using T = AdvTile<N, H>;
    const size_t smem = AdvSmem<N, UNIFORM, H>::BYTES + (HINT ? (size_t)T::NW * 32 * ADV_HINT_ROW : 0);     // + the per-warp classification bytes
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_advect_tile<N, SCHEME, UNIFORM, AFFINE, HINT, INTERP, BKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    AdvTmaMaps maps;
    const int tma_mask = build_advect_tma<N, H>(g, V, maps);
    const dim3 grd((g.n[0] + T::TX - 1) / T::TX, (g.n[1] + T::TY - 1) / T::TY, N == 3 ? (g.n[2] + T::TZ - 1) / T::TZ : 1);
    k_advect_tile<N, SCHEME, UNIFORM, AFFINE, HINT, INTERP, BKT><<<grd, T::NW * 32, smem, st>>>(g, co, index, V, alpha, dt, maps.m[0], maps.m[1], maps.m[2], tma_mask, ho.ws, ho.flag);
    return cudaSuccess;
}
template <int N, int SCHEME, bool UNIFORM, int AFFINE, bool HINT, int INTERP = 0>
static cudaError_t launch_advect_tile_h(const JpGrid &g, cudaStream_t st, Ptr3 co, const uint8_t *index, CPtr3 V, double alpha, double dt, const AdvHandoff &ho) {
    // the bucketed state compiled in for the variants of the time loops (trilinear, range grids); see jp_advect_tile.cuh
    if constexpr (INTERP == 0 && UNIFORM)
        if (g.bucketed) return launch_advect_tile_k<N, SCHEME, UNIFORM, AFFINE, HINT, INTERP, true>(g, st, co, index, V, alpha, dt, ho);
    return launch_advect_tile_k<N, SCHEME, UNIFORM, AFFINE, HINT, INTERP, false>(g, st, co, index, V, alpha, dt, ho);
}
template <int N, int SCHEME, bool UNIFORM, int AFFINE>
static cudaError_t launch_advect_tile(const JpGrid &g, cudaStream_t st, Ptr3 co, const uint8_t *index, CPtr3 V, double alpha, double dt, const AdvHandoff &ho) {
    return ho.flag ? launch_advect_tile_h<N, SCHEME, UNIFORM, AFFINE, true>(g, st, co, index, V, alpha, dt, ho)
                   : launch_advect_tile_h<N, SCHEME, UNIFORM, AFFINE, false>(g, st, co, index, V, alpha, dt, ho);
}

// *hinted is set when the launch left the move plan's classification words (tiled kernel only)
template <int N, int SCHEME>
static cudaError_t launch_advect(const JpGrid &g, dim3 grd, dim3 blk, cudaStream_t st, Ptr3 co, const uint8_t *index, CPtr3 V, double alpha, double dt,
                                 const AdvHandoff &hint, bool *hinted) {
    *hinted = false;
    if (jp_standard_staggering(g)) {
        *hinted = hint.flag != nullptr;
        if (g.uniform && g.affine == 2) return launch_advect_tile<N, SCHEME, true, 2>(g, st, co, index, V, alpha, dt, hint);
        if (g.uniform && g.affine == 1) return launch_advect_tile<N, SCHEME, true, 1>(g, st, co, index, V, alpha, dt, hint);
        return g.uniform ? launch_advect_tile<N, SCHEME, true, 0>(g, st, co, index, V, alpha, dt, hint)
                         : launch_advect_tile<N, SCHEME, false, 0>(g, st, co, index, V, alpha, dt, hint);
    }
    if (g.fast) {
        if (g.uniform) k_advect<N, SCHEME, true, true><<<grd, blk, 0, st>>>(g, co, index, V, alpha, dt);
        else           k_advect<N, SCHEME, true, false><<<grd, blk, 0, st>>>(g, co, index, V, alpha, dt);
    } else             k_advect<N, SCHEME, false, false><<<grd, blk, 0, st>>>(g, co, index, V, alpha, dt);
    return cudaSuccess;
}

static int advect_impl(jp_ctx *ctx, const jp_particles *p, int32_t scheme, double alpha, const double *const *V, double dt, int region, void *stream);
extern "C" int jp_advect(jp_ctx *ctx, const jp_particles *p, int32_t scheme, double alpha, const double *const *V, double dt, void *stream) {
    return advect_impl(ctx, p, scheme, alpha, V, dt, JP_REGION_ALL, stream);
}
// advection! in two launches so that update_cell_halo! can overlap the bulk of it: JP_REGION_SHELL advects every
// brick holding a cell of the two outermost cell layers (all the halo exchange reads or rewrites), JP_REGION_INTERIOR the rest.
// The two calls together are one jp_advect (same kernel, same results, same hand-off); the caller runs the exchange
// between them on another stream and joins before move_particles!.
extern "C" int jp_advect_region(jp_ctx *ctx, const jp_particles *p, int32_t scheme, double alpha, const double *const *V, double dt,
                                int32_t region, void *stream) {
    if (region != JP_REGION_ALL && region != JP_REGION_SHELL && region != JP_REGION_INTERIOR) return jp_fail(JP_ERR_INVALID, "jp_advect_region: unknown region");
    return advect_impl(ctx, p, scheme, alpha, V, dt, region, stream);
}
static int advect_impl(jp_ctx *ctx, const jp_particles *p, int32_t scheme, double alpha, const double *const *V, double dt, int region, void *stream) {
    PREP("jp_advect");
    if (region == JP_REGION_INTERIOR && !ctx->adv_split) return jp_fail(JP_ERR_INVALID, "jp_advect_region: JP_REGION_INTERIOR without a preceding JP_REGION_SHELL call");
    if (region != JP_REGION_INTERIOR && ctx->adv_split) { ctx->adv_split = 0; return jp_fail(JP_ERR_INVALID, "jp_advect_region: JP_REGION_SHELL must be followed by JP_REGION_INTERIOR"); }
    const bool tiled = jp_standard_staggering(g);
    if (region == JP_REGION_INTERIOR && !tiled) { ctx->adv_split = 0; return JP_OK; }      // the generic kernel did every cell in the shell call
    if (!V) return jp_fail(JP_ERR_INVALID, "jp_advect: null velocity tuple");
    CPtr3 v = {{nullptr, nullptr, nullptr}};
    for (int d = 0; d < g.ndim; d++) {
        if (!V[d]) return jp_fail(JP_ERR_INVALID, "jp_advect: null velocity component");
        v.p[d] = V[d];
    }
    if (scheme == JP_RK2 && !(0 < alpha && alpha < 1)) return jp_fail(JP_ERR_INVALID, "jp_advect: Only 0 < alpha < 1 is supported");
    if (scheme < 0 || scheme > 2) return jp_fail(JP_ERR_INVALID, "jp_advect: unknown integrator");
    if (region != JP_REGION_INTERIOR) hint_invalidate(ctx);          // the interior call completes the shell call's hand-off
    mi_invalidate(ctx);
    AdvHandoff hint;
    memset(&hint, 0, sizeof(hint));
    if (ctx->hint_opt && tiled && g.S <= JP_MAX_SLOTS) {
        int rc = move_plan_alloc(ctx);
        if (rc) return rc;
        hint.ws = ctx->mp; hint.flag = ctx->mp_flag;
        if (region != JP_REGION_INTERIOR) JP_CUDA(cudaMemsetAsync(ctx->mp_flag, 0, sizeof(unsigned int), st));
    }
    bool hinted = false;
    for (int ch = 0; ch < jp_nchunks(g); ch++) {          // one launch unless max_xcell > 64
        SlotChunk k = jp_chunk(g, ch);
        k.g.region = tiled ? region : 0;
        k.g.bucketed = region == JP_REGION_INTERIOR ? ctx->adv_split_bucketed : ctx->bucketed;
        const Ptr3 kc = jp_shift(co, k.off);
        const uint8_t *ki = p->index + k.off;
        cudaError_t le;
        if (g.ndim == 2) {
            if (scheme == 0) le = launch_advect<2, 0>(k.g, grd, blk, st, kc, ki, v, alpha, dt, hint, &hinted);
            else if (scheme == 1) le = launch_advect<2, 1>(k.g, grd, blk, st, kc, ki, v, alpha, dt, hint, &hinted);
            else le = launch_advect<2, 2>(k.g, grd, blk, st, kc, ki, v, alpha, dt, hint, &hinted);
        } else {
            if (scheme == 0) le = launch_advect<3, 0>(k.g, grd, blk, st, kc, ki, v, alpha, dt, hint, &hinted);
            else if (scheme == 1) le = launch_advect<3, 1>(k.g, grd, blk, st, kc, ki, v, alpha, dt, hint, &hinted);
            else le = launch_advect<3, 2>(k.g, grd, blk, st, kc, ki, v, alpha, dt, hint, &hinted);
        }
        if (le != cudaSuccess) return jp_fail(JP_ERR_CUDA, "jp_advect: %s", cudaGetErrorString(le));
    }
    JP_CHECK_LAUNCH();
    const bool hint_void = ctx->adv_split == 2;
    ctx->adv_split = region == JP_REGION_SHELL;
    if (region == JP_REGION_SHELL) ctx->adv_split_bucketed = ctx->bucketed;
    else ctx->bucketed = 0;             // (the interior launch of a split advection still sees the shell call's state)
    if (hinted && region != JP_REGION_SHELL && !hint_void) {               // after a shell call the words are still incomplete
        ctx->hint_valid = 1;
        for (int d = 0; d < 3; d++) ctx->hint_key[d] = d < g.ndim ? (const void *)p->coords[d] : nullptr;
        ctx->hint_key[3] = p->index;
    }
    return JP_OK;
}

// advection_LinP! / advection_MQS!: the tiled kernel (stencils idx - 1 .. idx + 2 served from a tile with one more halo node) on the
// standard staggering, the thread-per-cell kernel on global memory otherwise.  Same results either way (tests/test_advection_interpolants.py).
template <int N, int INTERP>
static cudaError_t launch_advect_hi(const JpGrid &g, dim3 grd, dim3 blk, cudaStream_t st, Ptr3 co, const uint8_t *index, CPtr3 V, int scheme, double alpha, double dt) {
    static const bool no_tile = getenv("JP_ADVECT_HI_GLOBAL") != nullptr;          // developer A/B switch
    if (!no_tile && jp_standard_staggering(g)) {
        const AdvHandoff none = {};
        JpGrid gt = g; gt.region = 0;        // (gt.bucketed comes with g)
        if (g.uniform) {
            if (scheme == 0) return launch_advect_tile_h<N, 0, true, 0, false, INTERP>(gt, st, co, index, V, alpha, dt, none);
            if (scheme == 1) return launch_advect_tile_h<N, 1, true, 0, false, INTERP>(gt, st, co, index, V, alpha, dt, none);
            return launch_advect_tile_h<N, 2, true, 0, false, INTERP>(gt, st, co, index, V, alpha, dt, none);
        }
        if (scheme == 0) return launch_advect_tile_h<N, 0, false, 0, false, INTERP>(gt, st, co, index, V, alpha, dt, none);
        if (scheme == 1) return launch_advect_tile_h<N, 1, false, 0, false, INTERP>(gt, st, co, index, V, alpha, dt, none);
        return launch_advect_tile_h<N, 2, false, 0, false, INTERP>(gt, st, co, index, V, alpha, dt, none);
    }
    if (scheme == 0) k_advect_hi<N, 0, INTERP><<<grd, blk, 0, st>>>(g, co, index, V, alpha, dt);
    else if (scheme == 1) k_advect_hi<N, 1, INTERP><<<grd, blk, 0, st>>>(g, co, index, V, alpha, dt);
    else k_advect_hi<N, 2, INTERP><<<grd, blk, 0, st>>>(g, co, index, V, alpha, dt);
    return cudaSuccess;
}

extern "C" int jp_advect_interp(jp_ctx *ctx, const jp_particles *p, int32_t scheme, double alpha, const double *const *V, double dt,
                                int32_t interp, void *stream) {
    if (interp == JP_INTERP_LINEAR) return jp_advect(ctx, p, scheme, alpha, V, dt, stream);
    PREP("jp_advect_interp");
    handoffs_invalidate(ctx);
    if (interp != JP_INTERP_LINP && interp != JP_INTERP_MQS) return jp_fail(JP_ERR_INVALID, "jp_advect_interp: unknown interpolant");
    if (!V) return jp_fail(JP_ERR_INVALID, "jp_advect_interp: null velocity tuple");
    CPtr3 v = {{nullptr, nullptr, nullptr}};
    for (int d = 0; d < g.ndim; d++) {
        if (!V[d]) return jp_fail(JP_ERR_INVALID, "jp_advect_interp: null velocity component");
        v.p[d] = V[d];
    }
    if (scheme == JP_RK2 && !(0 < alpha && alpha < 1)) return jp_fail(JP_ERR_INVALID, "jp_advect_interp: Only 0 < alpha < 1 is supported");
    if (scheme < 0 || scheme > 2) return jp_fail(JP_ERR_INVALID, "jp_advect_interp: unknown integrator");
    for (int ch = 0; ch < jp_nchunks(g); ch++) {
        SlotChunk k = jp_chunk(g, ch);
        k.g.bucketed = ctx->bucketed;
        const Ptr3 kc = jp_shift(co, k.off);
        const uint8_t *ki = p->index + k.off;
        cudaError_t le;
        if (g.ndim == 2) {
            if (interp == JP_INTERP_LINP) le = launch_advect_hi<2, 1>(k.g, grd, blk, st, kc, ki, v, scheme, alpha, dt);
            else                          le = launch_advect_hi<2, 2>(k.g, grd, blk, st, kc, ki, v, scheme, alpha, dt);
        } else {
            if (interp == JP_INTERP_LINP) le = launch_advect_hi<3, 1>(k.g, grd, blk, st, kc, ki, v, scheme, alpha, dt);
            else                          le = launch_advect_hi<3, 2>(k.g, grd, blk, st, kc, ki, v, scheme, alpha, dt);
        }
        if (le != cudaSuccess) return jp_fail(JP_ERR_CUDA, "jp_advect_interp: %s", cudaGetErrorString(le));
    }
    ctx->bucketed = 0;
    JP_CHECK_LAUNCH();
    return JP_OK;
}

static void move_plan_free(jp_ctx *ctx) {
    cudaFree(ctx->mp.code); cudaFree(ctx->mp.res); cudaFree(ctx->mp.occ0); cudaFree(ctx->mp.arrmask); cudaFree(ctx->mp.cnt); cudaFree(ctx->mp.off);
    cudaFree(ctx->mp_flag); cudaFree(ctx->cub_tmp);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->m_event) cudaEventDestroy(ctx->m_event);
    memset(&ctx->mp, 0, sizeof(ctx->mp));
    ctx->mp_flag = nullptr; ctx->cub_tmp = nullptr; ctx->cub_tmp_bytes = 0; ctx->h_pinned = nullptr; ctx->m_event = nullptr;
    ctx->mp_ready = 0;
}
// the plan workspace (~11 GB at 256^3): all or nothing -- a failed allocation frees what it got, so that the next call
// retries cleanly instead of launching kernels on a half-built workspace
static int move_plan_alloc(jp_ctx *ctx) {
    const JpGrid &g = ctx->g;
    if (ctx->mp_ready) return JP_OK;
    JP_NO_ALLOC_IN_CAPTURE("move_particles!: the plan workspace");
    cudaError_t e = cudaMalloc(&ctx->mp.code, sizeof(uint64_t) * (size_t)((g.S + 7) / 8) * g.C);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->mp.res, sizeof(uint64_t) * (size_t)((g.S + 7) / 8) * g.C);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->mp.occ0, sizeof(uint64_t) * g.C);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->mp.arrmask, sizeof(uint64_t) * g.C);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->mp.cnt, sizeof(uint32_t) * (g.C + 1));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->mp.off, sizeof(uint32_t) * (g.C + 1));
    if (e == cudaSuccess) e = cudaMemset(ctx->mp.cnt, 0, sizeof(uint32_t) * (g.C + 1));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->mp_flag, 2 * sizeof(unsigned int));          // [0] complex flag, [1] arrival count of the last call
    if (e == cudaSuccess) e = cudaMemset(ctx->mp_flag, 0, 2 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_pinned, 4 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->m_event, cudaEventDisableTiming);
    size_t tmp = 0;
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->mp.cnt, ctx->mp.off, (int)(g.C + 1));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->cub_tmp, tmp);
    if (e != cudaSuccess) {
        move_plan_free(ctx);
        cudaGetLastError();
        return jp_fail(JP_ERR_CUDA, "move_particles!: plan workspace: %s", cudaGetErrorString(e));
    }
    ctx->cub_tmp_bytes = tmp;
    ctx->mp.occ = ctx->occ; ctx->mp.leave = ctx->leave;
    ctx->mp_ready = 1;
    return JP_OK;
}

static int stage_reserve(jp_ctx *ctx, size_t elems) {
    if (elems <= ctx->stage_elems) return JP_OK;
    JP_NO_ALLOC_IN_CAPTURE("move_particles!: the staging buffer");
    if (ctx->stage) JP_CUDA(cudaFree(ctx->stage));
    ctx->stage = nullptr; ctx->stage_elems = 0;
    JP_CUDA(cudaMalloc(&ctx->stage, elems * sizeof(double)));
    ctx->stage_elems = elems;
    return JP_OK;
}

static int p2g_ws_reserve(jp_ctx *ctx) {
    const int NQ = ctx->g.ndim == 2 ? 4 : 8;
    if (!ctx->p2g_ws) {
        JP_NO_ALLOC_IN_CAPTURE("particle2grid!: the per-cell partial sums");
        JP_CUDA(cudaMalloc(&ctx->p2g_ws, sizeof(double) * 2 * NQ * ctx->g.C));
    }
    return JP_OK;
}

template <int N>
static void launch_scatter_interp(const JpGrid &g, dim3 grd, dim3 blk, cudaStream_t st, const MovePlanWs &ws, const MoveArrays &arrs, uint8_t *index,
                                  const double *stage, const MoveInterp &mi, bool fastw, const unsigned int *flag) {
    // the usual argument order (Fp first, phases second): 4-slot batches, compile-time array indices (jp_move_interp.cuh)
    static const bool no_fast = getenv("JP_SCI_GENERIC") != nullptr;           // developer A/B switch
    if (!no_fast && mi.iT == N && (mi.iP == N + 1 || mi.iP < 0) && arrs.n >= (mi.iP >= 0 ? N + 2 : N + 1)) {
#define JP_SCI_LAUNCH(KM, FW, PH) k_move_scatter_interp_fast<N, KM, FW, PH><<<grd, blk, 0, st>>>(g, ws, arrs, index, stage, mi, flag)
        if (mi.iP < 0) { if (fastw) JP_SCI_LAUNCH(2, true, false); else JP_SCI_LAUNCH(2, false, false); }
        else if (mi.K <= 2) { if (fastw) JP_SCI_LAUNCH(2, true, true); else JP_SCI_LAUNCH(2, false, true); }
        else { if (fastw) JP_SCI_LAUNCH(4, true, true); else JP_SCI_LAUNCH(4, false, true); }
#undef JP_SCI_LAUNCH
        return;
    }
    if (mi.K <= 2) {
        if (fastw) k_move_scatter_interp<N, 2, true><<<grd, blk, 0, st>>>(g, ws, arrs, index, stage, mi, flag);
        else       k_move_scatter_interp<N, 2, false><<<grd, blk, 0, st>>>(g, ws, arrs, index, stage, mi, flag);
    } else {
        if (fastw) k_move_scatter_interp<N, 4, true><<<grd, blk, 0, st>>>(g, ws, arrs, index, stage, mi, flag);
        else       k_move_scatter_interp<N, 4, false><<<grd, blk, 0, st>>>(g, ws, arrs, index, stage, mi, flag);
    }
}

// plan / gather / scatter path.  Everything is enqueued without waiting for the device: whether the call has to take the
// direct sweeps instead (a particle on a cell face, a move of more than one cell, a staging buffer that turned out too
// small) is decided ON THE DEVICE -- every kernel of this path returns at once when the flag is set, and the direct-sweep
// kernels enqueued behind it (jp_move) run only then.
// co-resident blocks of a cooperative kernel on this device (blocks of 256 threads, no dynamic shared memory)
template <typename K>
static int coop_max_blocks(jp_ctx *ctx, K kernel, int *cache) {
    if (*cache > 0) return JP_OK;
    int per_sm = 0, sms = 0, coop = 0;
    JP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) return jp_fail(JP_ERR_UNSUPPORTED, "this device does not support cooperative launches");
    JP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0));
    JP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    if (per_sm < 1) return jp_fail(JP_ERR_CUDA, "cooperative kernel does not fit an SM");
    *cache = per_sm * sms;
    return JP_OK;
}

template <int N>
static int move_planned(jp_ctx *ctx, const jp_particles *p, const JpArgs &a, cudaStream_t st) {
    const JpGrid &g = ctx->g;
    int rc = move_plan_alloc(ctx);
    if (rc) return rc;
    const dim3 blk(JP_BX, JP_BY, 1);
    const dim3 grd = tile_grid(g.n[0], g.n[1], g.n[2]);
    CPtr3 cco = {{p->coords[0], p->coords[1], p->coords[2]}};
    // JP_OPT_PROFILE: events between the stages (no synchronisation; jp_profile_read turns them into milliseconds later)
    cudaEvent_t *pev = nullptr;
    if (ctx->prof_opt && !ctx->capturing) {
        if (!ctx->prof_ev) {
            ctx->prof_ev = (cudaEvent_t *)calloc(JP_PROF_RING * JP_PROF_MARKS, sizeof(cudaEvent_t));
            for (int i = 0; i < JP_PROF_RING * JP_PROF_MARKS; i++) JP_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
        }
        pev = ctx->prof_ev + (size_t)(ctx->prof_calls % JP_PROF_RING) * JP_PROF_MARKS;
        ctx->prof_calls++;
    }
    int nev = 0;
    auto mark = [&]() { if (pev && nev < JP_PROF_MARKS) cudaEventRecord(pev[nev++], st); };
    mark();
    unsigned int *flag = ctx->mp_flag;
    const JpBox whole = {{0, 0, 0}, {g.n[0], g.n[1], g.n[2]}};
    bool use_hint = ctx->hint_valid && ctx->hint_key[3] == (const void *)p->index;
    for (int d = 0; d < N; d++) use_hint = use_hint && ctx->hint_key[d] == (const void *)p->coords[d];
    ctx->last_classify = use_hint ? 1 : 0;
    if (use_hint) {
        // the words (and the flag) were left by jp_advect; only planes rewritten by jp_halo_unpack are redone
        for (int i = 0; i < ctx->hint_ndirty; i++) {
            JpBox bx = whole;
            bx.o[ctx->hint_dirty[i][0]] = ctx->hint_dirty[i][1];
            bx.e[ctx->hint_dirty[i][0]] = 1;
            k_move_classify3<N, true><<<tile_grid(bx.e[0], bx.e[1], N == 3 ? bx.e[2] : 1), blk, 0, st>>>(g, cco, p->index, ctx->mp, flag, bx);
        }
    } else {
        JP_CUDA(cudaMemsetAsync(flag, 0, sizeof(unsigned int), st));
        k_move_classify3<N, false><<<grd, blk, 0, st>>>(g, cco, p->index, ctx->mp, flag, whole);
    }
    hint_invalidate(ctx);
    JP_CHECK_LAUNCH();
    mark();
    MoveArrays arrs; arrs.n = 0;
    for (int d = 0; d < N; d++) arrs.a[arrs.n++] = p->coords[d];
    for (int i = 0; i < a.n; i++) arrs.a[arrs.n++] = a.a[i];
    const size_t AS = (size_t)((arrs.n + 3) & ~3);              // array-of-structs staging, stride padded to 4 doubles
    // staging capacity: the arrival count of the previous call comes back asynchronously (pinned word + event); grow when it is
    // known and calls for it.  Never blocks except to reallocate.
    // (While the stream is captured into a CUDA graph neither the query nor the read-back below happen: the replays run with the
    // buffer the eager calls before the capture sized; one that turns out too small sends that replay to the direct sweeps.)
    if (!ctx->capturing && ctx->m_pending && cudaEventQuery(ctx->m_event) == cudaSuccess) {
        ctx->m_pending = 0;
        const size_t lastM = ctx->h_pinned[1];
        if ((lastM + lastM / 4 + 2048) * AS > ctx->stage_elems) {           // keep >= 25 % + 2048 rows of head room over the last count ...
            rc = stage_reserve(ctx, (lastM + lastM / 2 + 4096) * AS);      // ... by growing to 50 % + 4096 (grow-only)
            if (rc) return rc;
        }
    }
    const int ncx = (g.n[0] + 2) / 3, ncy = (g.n[1] + 2) / 3, ncz = N == 3 ? (g.n[2] + 2) / 3 : 1;
    const int64_t ncol = (int64_t)ncx * ncy * ncz;
    const unsigned nblk = (unsigned)((ncol + 255) / 256);
    const unsigned cblk = (unsigned)((g.C + 255) / 256);
    if (ctx->move_policy == JP_MOVE_POLICY_DENSE) k_move_prevacate<<<cblk, 256, 0, st>>>(g.C, ctx->mp, flag);
    rc = coop_max_blocks(ctx, k_move_plan_all<N>, &ctx->coop_plan_max);
    if (rc) return rc;
    if (nblk > (unsigned)ctx->coop_plan_max) {               // a colour does not fit the device at once: one launch per colour
        for (int ox = 0; ox < 3; ox++)
            for (int oy = 0; oy < 3; oy++)
                for (int oz = 0; oz < (N == 3 ? 3 : 1); oz++)
                    k_move_plan<N><<<nblk, 256, 0, st>>>(g, ctx->mp, ox, oy, oz, ncx, ncy, ncol, ctx->stats, ctx->move_policy, flag);
    } else {                                                 // the 3^N ordered colours in one cooperative launch
        JpGrid gk = g; MovePlanWs wk = ctx->mp; int kx = ncx, ky = ncy, pol = ctx->move_policy; int64_t kn = ncol; long long *ks = ctx->stats;
        const unsigned int *kf = flag;
        void *kargs[] = {&gk, &wk, &kx, &ky, &kn, &ks, &pol, &kf};
        JP_CUDA(cudaLaunchCooperativeKernel((const void *)k_move_plan_all<N>, dim3(nblk), dim3(256), kargs, 0, st));
    }
    mark();
    k_move_finalize<N><<<cblk, 256, 0, st>>>(g, ctx->mp, flag);
    JP_CHECK_LAUNCH();
    JP_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, ctx->cub_tmp_bytes, ctx->mp.cnt, ctx->mp.off, (int)(g.C + 1), st));
    if (!ctx->m_probed && ctx->capturing)
        return jp_fail(JP_ERR_INVALID, "move_particles!: the first call on a context sizes the staging buffer with a read-back: run the step once eagerly before capturing it into a CUDA graph");
    if (!ctx->m_probed) {
        // very first planned call on this context: learn the arrival count now (the only blocking read-back there is)
        ctx->m_probed = 1;
        JP_CUDA(cudaMemcpyAsync(ctx->h_pinned + 1, ctx->mp.off + g.C, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        JP_CUDA(cudaStreamSynchronize(st));
        const size_t M0 = ctx->h_pinned[1];
        rc = stage_reserve(ctx, (M0 + M0 / 2 + 4096) * AS);
        if (rc) return rc;
    }
    k_move_after_scan<<<1, 1, 0, st>>>(ctx->stats, ctx->mp.off + g.C, (uint64_t)(ctx->stage_elems / AS), flag, flag + 1);
    if (!ctx->m_pending && !ctx->capturing) {
        JP_CUDA(cudaMemcpyAsync(ctx->h_pinned + 1, flag + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        JP_CUDA(cudaEventRecord(ctx->m_event, st));
        ctx->m_pending = 1;
    }
    mark();
    k_move_gather<N><<<grd, blk, 0, st>>>(g, ctx->mp, arrs, ctx->stage, flag);
    mark();
    // move -> interpolation hand-off: the scatter also leaves particle2grid!'s cell sums / the centre phase ratios
    MoveInterp mi; mi.iT = mi.iP = -1; mi.K = 0; mi.PW = mi.PWF = mi.RC = nullptr;
    if (ctx->mi_opt) {
        const bool twopass = ctx->p2g_mode == JP_P2G_TWOPASS || ctx->p2g_mode == JP_P2G_TWOPASS_FASTW;
        for (int i = 0; i < a.n; i++) {
            if (twopass && ctx->mi_fp && a.a[i] == ctx->mi_fp) mi.iT = N + i;
            if (ctx->mi_ph && a.a[i] == ctx->mi_ph && ctx->mi_K >= 1 && ctx->mi_K <= 4) mi.iP = N + i;
        }
        if (mi.iT >= 0) {
            rc = p2g_ws_reserve(ctx); if (rc) return rc;
            mi.PW = ctx->p2g_ws; mi.PWF = ctx->p2g_ws + (int64_t)(N == 2 ? 4 : 8) * g.C;
        }
        if (mi.iP >= 0) {
            mi.K = ctx->mi_K;
            const size_t need = (size_t)mi.K * g.C;
            if (need > ctx->mi_rc_elems) {
                JP_NO_ALLOC_IN_CAPTURE("move -> interpolation hand-off: the centre-ratio workspace");
                if (ctx->mi_rc) JP_CUDA(cudaFree(ctx->mi_rc));
                ctx->mi_rc = nullptr; ctx->mi_rc_elems = 0;
                JP_CUDA(cudaMalloc(&ctx->mi_rc, need * sizeof(double)));
                ctx->mi_rc_elems = need;
            }
            mi.RC = ctx->mi_rc;
        }
    }
    if (mi.iT >= 0 || mi.iP >= 0) {
        launch_scatter_interp<N>(g, grd, blk, st, ctx->mp, arrs, p->index, ctx->stage, mi, ctx->p2g_mode == JP_P2G_TWOPASS_FASTW, flag);
        ctx->mi_valid_p2g = mi.iT >= 0; ctx->mi_valid_ph = mi.iP >= 0; ctx->mi_p2g_mode = ctx->p2g_mode;
        for (int d = 0; d < 3; d++) ctx->mi_key[d] = d < N ? (const void *)p->coords[d] : nullptr;
        ctx->mi_key[3] = p->index;
    } else
        k_move_scatter<N><<<grd, blk, 0, st>>>(g, ctx->mp, arrs, p->index, ctx->stage, flag);
    mark();
    JP_CHECK_LAUNCH();
    return JP_OK;
}

// mean duration (ms) of each stage of the planned jp_move over the calls recorded since the last read:
// out[0..4] = classify, plan (3^N launches), finalize + scan, gather, scatter(+interp).  Synchronises the device.
extern "C" int jp_profile_read(jp_ctx *ctx, double *out_ms, int32_t *ncalls) {
    if (!ctx || !out_ms) return jp_fail(JP_ERR_INVALID, "jp_profile_read: null argument");
    JP_CUDA(cudaSetDevice(ctx->device));
    JP_CUDA(cudaDeviceSynchronize());
    const int n = ctx->prof_calls < JP_PROF_RING ? ctx->prof_calls : JP_PROF_RING;
    for (int k = 0; k < JP_PROF_MARKS - 1; k++) out_ms[k] = 0.0;
    for (int c = 0; c < n; c++)
        for (int k = 0; k < JP_PROF_MARKS - 1; k++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->prof_ev[c * JP_PROF_MARKS + k], ctx->prof_ev[c * JP_PROF_MARKS + k + 1]);
            out_ms[k] += ms / n;
        }
    if (ncalls) *ncalls = n;
    ctx->prof_calls = 0;
    cudaGetLastError();
    return JP_OK;
}

extern "C" int jp_move(jp_ctx *ctx, const jp_particles *p, double *const *args, int32_t nargs, void *stream) {
    PREP("jp_move");
    JpArgs a;
    int rc = pack_args(args, nargs, a, "jp_move");
    if (rc) return rc;
    mi_invalidate(ctx);
    ctx->bucketed = 1;                                       // whatever path the call takes, it leaves every particle inside its cell
    ctx->last_stream = stream;
    JP_CUDA(cudaMemsetAsync(ctx->stats, 0, 3 * sizeof(long long), st));
    if (ctx->move_policy == JP_MOVE_POLICY_DENSE && (g.S > JP_MAX_SLOTS || ctx->move_mode != JP_MOVE_AUTO))
        return jp_fail(JP_ERR_INVALID, "jp_move: JP_MOVE_POLICY_DENSE needs the planned path (%s)", "JP_MOVE_AUTO, max_xcell <= 64");
    // the in-place sweeps cannot vacate first (the slots still hold the leavers' payload): a DENSE call the planner declines
    // (a particle on a face, a far move) runs them with the reference's rule -- jp_last_move_path tells
    const int sweep_policy = ctx->move_policy == JP_MOVE_POLICY_COMPACT;
    if (g.S > JP_MAX_SLOTS) {                                // wide cells: literal per-cell sweeps on the index bytes
        ctx->last_move_path = 1;
        hint_invalidate(ctx);
        const int ncx = (g.n[0] + 2) / 3, ncy = (g.n[1] + 2) / 3, ncz = g.ndim == 3 ? (g.n[2] + 2) / 3 : 1;
        const int64_t ncol = (int64_t)ncx * ncy * ncz;
        const unsigned nblk = (unsigned)((ncol + 127) / 128);
        for (int ox = 0; ox < 3; ox++)
            for (int oy = 0; oy < 3; oy++)
                for (int oz = 0; oz < (g.ndim == 3 ? 3 : 1); oz++) {
                    if (g.ndim == 2) k_move_sweep_wide<2><<<nblk, 128, 0, st>>>(g, co, p->index, a, ox, oy, oz, ncx, ncy, ncol, ctx->stats, sweep_policy);
                    else             k_move_sweep_wide<3><<<nblk, 128, 0, st>>>(g, co, p->index, a, ox, oy, oz, ncx, ncy, ncol, ctx->stats, sweep_policy);
                }
        JP_CHECK_LAUNCH();
        return JP_OK;
    }
    const unsigned int *run_flag = nullptr;                  // JP_MOVE_DIRECT: the sweeps always run
    if (ctx->move_mode == JP_MOVE_AUTO) {
        rc = g.ndim == 2 ? move_planned<2>(ctx, p, a, st) : move_planned<3>(ctx, p, a, st);
        if (rc) return rc;
        ctx->last_move_path = -1;                            // decided on the device: jp_last_move_path reads the flag
        run_flag = ctx->mp_flag;                             // the sweeps below run only if the planned path declined
    } else ctx->last_move_path = 1;
    hint_invalidate(ctx);
    if (g.ndim == 2) k_move_classify<2><<<grd, blk, 0, st>>>(g, cco, p->index, ctx->occ, ctx->leave, run_flag);
    else             k_move_classify<3><<<grd, blk, 0, st>>>(g, cco, p->index, ctx->occ, ctx->leave, run_flag);
    JP_CHECK_LAUNCH();
    const int ncx = (g.n[0] + 2) / 3, ncy = (g.n[1] + 2) / 3, ncz = g.ndim == 3 ? (g.n[2] + 2) / 3 : 1;
    const int64_t ncol = (int64_t)ncx * ncy * ncz;          // source cells per colour (upper bound)
    const int64_t want_blk = (ncol + 7) / 8;                 // one warp per source cell, 8 warps per block, grid-stride beyond what is co-resident
    {
        JpGrid gk = g; Ptr3 kco = co; uint8_t *kidx = p->index; JpArgs ka = a; uint64_t *kocc = ctx->occ, *klv = ctx->leave;
        int kx = ncx, ky = ncy, pol = sweep_policy; int64_t kn = ncol; long long *ks = ctx->stats; const unsigned int *kf = run_flag;
        void *kargs[] = {&gk, &kco, &kidx, &ka, &kocc, &klv, &kx, &ky, &kn, &ks, &pol, &kf};
        if (g.ndim == 2) rc = coop_max_blocks(ctx, k_move_sweep_all<2>, &ctx->coop_sweep_max);
        else             rc = coop_max_blocks(ctx, k_move_sweep_all<3>, &ctx->coop_sweep_max);
        if (rc) return rc;
        const unsigned nblk = (unsigned)(want_blk < ctx->coop_sweep_max ? want_blk : ctx->coop_sweep_max);
        JP_CUDA(cudaLaunchCooperativeKernel(g.ndim == 2 ? (const void *)k_move_sweep_all<2> : (const void *)k_move_sweep_all<3>, dim3(nblk), dim3(256), kargs, 0, st));
    }
    JP_CHECK_LAUNCH();
    return JP_OK;
}

// bits 0-7: 0 = the last jp_move took the plan / gather / scatter path, 1 = direct sweeps; bits 8+: why (1 far move, 2 particle
// on a face of its own cell, 4 on a face of its destination, 8 staging buffer too small).  In JP_MOVE_AUTO the choice was made
// on the device, so this reads the flag back (synchronises the stream of that jp_move).
extern "C" int jp_last_move_path(const jp_ctx *cctx) {
    jp_ctx *ctx = const_cast<jp_ctx *>(cctx);
    if (!ctx) return -1;
    if (ctx->last_move_path >= 0) return ctx->last_move_path;
    unsigned int f = 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(&f, ctx->mp_flag, sizeof(f), cudaMemcpyDeviceToHost, (cudaStream_t)ctx->last_stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize((cudaStream_t)ctx->last_stream) != cudaSuccess) return -1;
    ctx->last_complex = (int)f;
    return (f ? 1 : 0) | ((int)f << 8);
}

extern "C" int jp_move_stats(jp_ctx *ctx, int64_t out[3], void *stream) {
    if (!ctx || !out) return jp_fail(JP_ERR_INVALID, "jp_move_stats: null argument");
    JP_CUDA(cudaSetDevice(ctx->device));
    long long h[3];
    JP_CUDA(cudaMemcpyAsync(h, ctx->stats, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    JP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
    return JP_OK;
}

// wide cells (max_xcell > 64): 2^N colour sweeps of the literal per-cell routine, colour order of the reference
template <bool PHASE>
static void launch_inject_wide(const JpGrid &g, cudaStream_t st, Ptr3 co, uint8_t *index, const JpArgs &a, int min_xcell, uint64_t seed, uint32_t step,
                               long long *stats, const InjPhase &ph) {
    const int ncx = (g.n[0] + 1) / 2, ncy = (g.n[1] + 1) / 2, ncz = g.ndim == 3 ? (g.n[2] + 1) / 2 : 1;
    const int64_t ncol = (int64_t)ncx * ncy * ncz;
    const unsigned nblk = (unsigned)((ncol + 127) / 128);
    for (int ox = 0; ox < 2; ox++)
        for (int oy = 0; oy < 2; oy++)
            for (int oz = 0; oz < (g.ndim == 3 ? 2 : 1); oz++) {
                if (g.ndim == 2) k_inject_wide<2, PHASE><<<nblk, 128, 0, st>>>(g, co, index, a, ox, oy, oz, ncx, ncy, ncol, min_xcell, seed, step, stats, ph);
                else             k_inject_wide<3, PHASE><<<nblk, 128, 0, st>>>(g, co, index, a, ox, oy, oz, ncx, ncy, ncol, min_xcell, seed, step, stats, ph);
            }
}

template <int N, bool PHASE>
static int launch_inject_sweeps(jp_ctx *ctx, cudaStream_t st, Ptr3 co, uint8_t *index, const JpArgs &a, int min_xcell, uint64_t seed, uint32_t step,
                                const InjPhase &ph) {
    for (int col = 0; col < (N == 3 ? 8 : 4); col++)
        k_inject_sweep<N, PHASE><<<148 * JP_MINB_INJECT * 2, 256, 0, st>>>(ctx->g, co, index, a, ctx->occ, ctx->inbox, ctx->inj_list + (int64_t)col * ctx->inj_cap,
                                                                          ctx->inj_count + col, min_xcell, seed, step, ctx->stats, ph);
    return JP_OK;
}

extern "C" int jp_inject(jp_ctx *ctx, const jp_particles *p, double *const *args, int32_t nargs, int32_t min_xcell, uint64_t seed, uint32_t step, void *stream) {
    PREP("jp_inject");
    handoffs_invalidate(ctx);
    JpArgs a;
    int rc = pack_args(args, nargs, a, "jp_inject");
    if (rc) return rc;
    if (step >= (1u << 31)) return jp_fail(JP_ERR_INVALID, "jp_inject: step must be < 2^31");
    const int NQ = g.ndim == 2 ? 4 : 8;
    const int min_xq = (min_xcell + NQ - 1) / NQ;
    JP_CUDA(cudaMemsetAsync(ctx->stats + 3, 0, sizeof(long long), st));
    if ((int64_t)g.C >= (1ll << 31)) return jp_fail(JP_ERR_UNSUPPORTED, "jp_inject: more than 2^31 cells");
    if (g.S > JP_MAX_SLOTS) {
        launch_inject_wide<false>(g, st, co, p->index, a, min_xcell, seed, step, ctx->stats, InjPhase());
        if (ctx->capturing) k_step_advance<<<1, 1, 0, st>>>(ctx->stats);
        JP_CHECK_LAUNCH();
        return JP_OK;
    }
    JP_CUDA(cudaMemsetAsync(ctx->inj_count, 0, 8 * sizeof(unsigned int), st));
    if (g.ndim == 2) k_inject_classify<2, false><<<grd, blk, 0, st>>>(g, cco, p->index, min_xq, ctx->occ, ctx->inbox, ctx->inj_list, ctx->inj_count, ctx->inj_cap);
    else             k_inject_classify<3, false><<<grd, blk, 0, st>>>(g, cco, p->index, min_xq, ctx->occ, ctx->inbox, ctx->inj_list, ctx->inj_count, ctx->inj_cap);
    JP_CHECK_LAUNCH();
    // colour order of the reference: offset_i outermost (src/Particles/injection.jl:30-49)
    rc = g.ndim == 2 ? launch_inject_sweeps<2, false>(ctx, st, co, p->index, a, min_xcell, seed, step, InjPhase())
                     : launch_inject_sweeps<3, false>(ctx, st, co, p->index, a, min_xcell, seed, step, InjPhase());
    if (rc) return rc;
    if (ctx->capturing) k_step_advance<<<1, 1, 0, st>>>(ctx->stats);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_inject_phase(jp_ctx *ctx, const jp_particles *p, double *phases, double *const *args, const double *const *fields,
                               const int32_t *field_kind, int32_t nargs, int32_t min_xcell, uint64_t seed, uint32_t step, void *stream) {
    PREP("jp_inject_phase");
    handoffs_invalidate(ctx);
    JpArgs a;
    int rc = pack_args(args, nargs, a, "jp_inject_phase");
    if (rc) return rc;
    if (!phases) return jp_fail(JP_ERR_INVALID, "jp_inject_phase: null phase array");
    if (step >= (1u << 31)) return jp_fail(JP_ERR_INVALID, "jp_inject_phase: step must be < 2^31");
    if ((int64_t)g.C >= (1ll << 31)) return jp_fail(JP_ERR_UNSUPPORTED, "jp_inject_phase: more than 2^31 cells");
    InjPhase ph;
    memset(&ph, 0, sizeof(ph));
    ph.phases = phases;
    for (int j = 0; j < nargs; j++) {
        if (!fields || !fields[j] || !field_kind) return jp_fail(JP_ERR_INVALID, "jp_inject_phase: null grid field");
        if (field_kind[j] != 0 && field_kind[j] != 1) return jp_fail(JP_ERR_INVALID, "jp_inject_phase: field kind must be 0 (vertex) or 1 (centre)");
        ph.fields[j] = fields[j]; ph.fkind[j] = field_kind[j];
    }
    const int NQ = g.ndim == 2 ? 4 : 8;
    const int min_xq = (min_xcell + NQ - 1) / NQ;
    JP_CUDA(cudaMemsetAsync(ctx->stats + 3, 0, sizeof(long long), st));
    if (g.S > JP_MAX_SLOTS) {
        launch_inject_wide<true>(g, st, co, p->index, a, min_xcell, seed, step, ctx->stats, ph);
        if (ctx->capturing) k_step_advance<<<1, 1, 0, st>>>(ctx->stats);
        JP_CHECK_LAUNCH();
        return JP_OK;
    }
    JP_CUDA(cudaMemsetAsync(ctx->inj_count, 0, 8 * sizeof(unsigned int), st));
    if (g.ndim == 2) k_inject_classify<2, true><<<grd, blk, 0, st>>>(g, cco, p->index, min_xq, ctx->occ, ctx->inbox, ctx->inj_list, ctx->inj_count, ctx->inj_cap);
    else             k_inject_classify<3, true><<<grd, blk, 0, st>>>(g, cco, p->index, min_xq, ctx->occ, ctx->inbox, ctx->inj_list, ctx->inj_count, ctx->inj_cap);
    JP_CHECK_LAUNCH();
    rc = g.ndim == 2 ? launch_inject_sweeps<2, true>(ctx, st, co, p->index, a, min_xcell, seed, step, ph)
                     : launch_inject_sweeps<3, true>(ctx, st, co, p->index, a, min_xcell, seed, step, ph);
    if (rc) return rc;
    if (ctx->capturing) k_step_advance<<<1, 1, 0, st>>>(ctx->stats);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_inject_stats(jp_ctx *ctx, int64_t *out, void *stream) {
    if (!ctx || !out) return jp_fail(JP_ERR_INVALID, "jp_inject_stats: null argument");
    JP_CUDA(cudaSetDevice(ctx->device));
    long long h;
    JP_CUDA(cudaMemcpyAsync(&h, ctx->stats + 3, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    JP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *out = h;
    return JP_OK;
}

extern "C" int jp_force_injection(jp_ctx *ctx, const jp_particles *p, const double *const *pnew, double *const *fields, const double *values,
                                  int32_t nfields, void *stream) {
    PREP("jp_force_injection");
    handoffs_invalidate(ctx);
    ctx->bucketed = 0;                                       // the caller's points go where the caller says
    JpArgs f;
    int rc = pack_args(fields, nfields, f, "jp_force_injection");
    if (rc) return rc;
    if (!pnew) return jp_fail(JP_ERR_INVALID, "jp_force_injection: null p_new");
    CPtr3 pn = {{nullptr, nullptr, nullptr}};
    for (int d = 0; d < g.ndim; d++) {
        if (!pnew[d]) return jp_fail(JP_ERR_INVALID, "jp_force_injection: null p_new component");
        pn.p[d] = pnew[d];
    }
    if (nfields > 0 && !values) return jp_fail(JP_ERR_INVALID, "jp_force_injection: null values");
    ForceVals fv;
    for (int a = 0; a < JP_MAX_ARGS; a++) fv.v[a] = a < nfields ? values[a] : 0.0;
    if (g.ndim == 2) k_force_injection<2><<<grd, blk, 0, st>>>(g, co, p->index, pn, f, fv);
    else             k_force_injection<3><<<grd, blk, 0, st>>>(g, co, p->index, pn, f, fv);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_clean(jp_ctx *ctx, const jp_particles *p, double *const *args, int32_t nargs, void *stream) {
    PREP("jp_clean");
    handoffs_invalidate(ctx);
    ctx->bucketed = 1;
    JpArgs a;
    int rc = pack_args(args, nargs, a, "jp_clean");
    if (rc) return rc;
    for (int ch = 0; ch < jp_nchunks(g); ch++) {
        const SlotChunk k = jp_chunk(g, ch);
        if (g.ndim == 2) k_clean<2><<<grd, blk, 0, st>>>(k.g, jp_shift(co, k.off), p->index + k.off, jp_shift(a, k.off));
        else             k_clean<3><<<grd, blk, 0, st>>>(k.g, jp_shift(co, k.off), p->index + k.off, jp_shift(a, k.off));
    }
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_grid2particle(jp_ctx *ctx, const jp_particles *p, double *Fp, const double *F, void *stream) {
    PREP("jp_grid2particle");
    mi_invalidate(ctx);                                  // a particle field is rewritten
    if (!Fp || !F) return jp_fail(JP_ERR_INVALID, "jp_grid2particle: null field");
    for (int ch = 0; ch < jp_nchunks(g); ch++) {
        const SlotChunk k = jp_chunk(g, ch);
        if (g.ndim == 2) k_g2p<2><<<grd, blk, 0, st>>>(k.g, jp_shift(cco, k.off), p->index + k.off, Fp + k.off, F);
        else             k_g2p<3><<<grd, blk, 0, st>>>(k.g, jp_shift(cco, k.off), p->index + k.off, Fp + k.off, F);
    }
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_grid2particle_flip(jp_ctx *ctx, const jp_particles *p, double *Fp, const double *F, const double *F0, double alpha, void *stream) {
    PREP("jp_grid2particle_flip");
    mi_invalidate(ctx);                                  // a particle field is rewritten
    if (!Fp || !F || !F0) return jp_fail(JP_ERR_INVALID, "jp_grid2particle_flip: null field");
    for (int ch = 0; ch < jp_nchunks(g); ch++) {
        const SlotChunk k = jp_chunk(g, ch);
        if (g.ndim == 2) k_g2p_flip<2><<<grd, blk, 0, st>>>(k.g, jp_shift(cco, k.off), p->index + k.off, Fp + k.off, F, F0, alpha);
        else             k_g2p_flip<3><<<grd, blk, 0, st>>>(k.g, jp_shift(cco, k.off), p->index + k.off, Fp + k.off, F, F0, alpha);
    }
    JP_CHECK_LAUNCH();
    return JP_OK;
}

// subgrid_diffusion!(pT, T_grid, dT_grid, subgrid_arrays, particles, dt; d) / subgrid_diffusion_centroid!
extern "C" int jp_subgrid_diffusion(jp_ctx *ctx, const jp_particles *p, double *pT, const double *T_grid, const double *dT_grid,
                                    const int32_t *dT_extents, double *pT0, double *pdT, const double *dt0, double *dT_subgrid,
                                    double dt, double d, int32_t centroid, void *stream) {
    PREP("jp_subgrid_diffusion");
    mi_invalidate(ctx);                                  // a particle field is rewritten
    if (!pT || !T_grid || !dT_grid || !dT_extents || !pT0 || !pdT || !dt0 || !dT_subgrid) return jp_fail(JP_ERR_INVALID, "jp_subgrid_diffusion: null argument");
    const int plus = centroid ? 0 : 1;
    const int n0 = g.n[0] + plus, n1 = g.n[1] + plus, n2 = g.ndim == 3 ? g.n[2] + plus : 1;
    for (int dd = 0; dd < g.ndim; dd++)
        if (dT_extents[dd] < (dd == 0 ? n0 : dd == 1 ? n1 : n2) + 1) return jp_fail(JP_ERR_INVALID, "jp_subgrid_diffusion: dT_grid needs one more node than dT_subgrid per dimension (it is read at I + 1)");
    if (g.ndim == 2) {
        if (centroid) k_subgrid_pass1<2, true><<<grd, blk, 0, st>>>(g, cco, p->index, pT, pT0, pdT, dt0, T_grid, d, dt);
        else          k_subgrid_pass1<2, false><<<grd, blk, 0, st>>>(g, cco, p->index, pT, pT0, pdT, dt0, T_grid, d, dt);
    } else {
        if (centroid) k_subgrid_pass1<3, true><<<grd, blk, 0, st>>>(g, cco, p->index, pT, pT0, pdT, dt0, T_grid, d, dt);
        else          k_subgrid_pass1<3, false><<<grd, blk, 0, st>>>(g, cco, p->index, pT, pT0, pdT, dt0, T_grid, d, dt);
    }
    JP_CHECK_LAUNCH();
    int rc = centroid ? jp_particle2centroid(ctx, p, dT_subgrid, pdT, stream) : jp_particle2grid(ctx, p, dT_subgrid, pdT, stream);
    if (rc) return rc;
    const dim3 ng = tile_grid(n0, n1, n2);
    if (g.ndim == 2) k_update_dT_subgrid<2><<<ng, blk, 0, st>>>(dT_subgrid, dT_grid, n0, n1, n2, dT_extents[0], dT_extents[1]);
    else             k_update_dT_subgrid<3><<<ng, blk, 0, st>>>(dT_subgrid, dT_grid, n0, n1, n2, dT_extents[0], dT_extents[1]);
    if (g.ndim == 2) {
        if (centroid) k_subgrid_pass2<2, true><<<grd, blk, 0, st>>>(g, cco, p->index, pT, pT0, pdT, dT_subgrid);
        else          k_subgrid_pass2<2, false><<<grd, blk, 0, st>>>(g, cco, p->index, pT, pT0, pdT, dT_subgrid);
    } else {
        if (centroid) k_subgrid_pass2<3, true><<<grd, blk, 0, st>>>(g, cco, p->index, pT, pT0, pdT, dT_subgrid);
        else          k_subgrid_pass2<3, false><<<grd, blk, 0, st>>>(g, cco, p->index, pT, pT0, pdT, dT_subgrid);
    }
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_centroid2particle(jp_ctx *ctx, const jp_particles *p, double *Fp, const double *Fc, void *stream) {
    PREP("jp_centroid2particle");
    mi_invalidate(ctx);                                  // a particle field is rewritten
    if (!Fp || !Fc) return jp_fail(JP_ERR_INVALID, "jp_centroid2particle: null field");
    if (g.ndim == 2) k_c2p<2><<<grd, blk, 0, st>>>(g, cco, Fp, Fc);
    else             k_c2p<3><<<grd, blk, 0, st>>>(g, cco, Fp, Fc);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_set_option(jp_ctx *ctx, int32_t option, int32_t value) {
    if (!ctx) return jp_fail(JP_ERR_INVALID, "jp_set_option: null context");
    if (option == JP_OPT_MOVE_MODE && (value == JP_MOVE_AUTO || value == JP_MOVE_DIRECT)) { ctx->move_mode = value; return JP_OK; }
    if (option == JP_OPT_P2G_MODE && (value == JP_P2G_EXACT || value == JP_P2G_TWOPASS || value == JP_P2G_TWOPASS_FASTW)) { ctx->p2g_mode = value; return JP_OK; }
    if (option == JP_OPT_MOVE_POLICY && (value == JP_MOVE_POLICY_REFERENCE || value == JP_MOVE_POLICY_COMPACT || value == JP_MOVE_POLICY_DENSE)) { ctx->move_policy = value; return JP_OK; }
    if (option == JP_OPT_ADVECT_AFFINE && (value == 0 || value == 1)) { ctx->g.affine = value ? ctx->affine_detected : 0; return JP_OK; }
    if (option == JP_OPT_ADVECT_CLASSIFY && (value == 0 || value == 1)) { ctx->hint_opt = value; hint_invalidate(ctx); return JP_OK; }
    if (option == JP_OPT_MOVE_INTERP && (value == 0 || value == 1)) { ctx->mi_opt = value; mi_invalidate(ctx); return JP_OK; }
    if (option == JP_OPT_GRAPH_STEP_OFFSET && value >= 0) {
        const unsigned int v = (unsigned int)value;
        JP_CUDA(cudaSetDevice(ctx->device));
        JP_CUDA(cudaMemcpy(jp_step_offset(ctx->stats), &v, sizeof(v), cudaMemcpyHostToDevice));
        return JP_OK;
    }
    if (option == JP_OPT_PROFILE && (value == 0 || value == 1)) { ctx->prof_opt = value; ctx->prof_calls = 0; return JP_OK; }
    return jp_fail(JP_ERR_INVALID, "jp_set_option: unknown option/value");
}

// The caller changed coordinates, the index mask or a registered particle field by other means than this library: drop what the
// hand-offs left (the next jp_move classifies from the coordinates, the next particle2grid! / phase_ratios_center! stream the particles).
extern "C" int jp_invalidate_handoffs(jp_ctx *ctx) {
    if (!ctx) return jp_fail(JP_ERR_INVALID, "jp_invalidate_handoffs: null context");
    handoffs_invalidate(ctx);
    ctx->bucketed = 0;
    return JP_OK;
}

extern "C" int jp_move_interp_fields(jp_ctx *ctx, const double *Fp, const double *phases, int32_t K) {
    if (!ctx) return jp_fail(JP_ERR_INVALID, "jp_move_interp_fields: null context");
    if (phases && (K < 1 || K > JP_MAX_PHASES)) return jp_fail(JP_ERR_UNSUPPORTED, "jp_move_interp_fields: 1 <= nphases <= 32 required");
    ctx->mi_fp = Fp; ctx->mi_ph = phases; ctx->mi_K = phases ? K : 0;
    mi_invalidate(ctx);
    return JP_OK;
}

extern "C" int jp_get_option(const jp_ctx *ctx, int32_t option, int32_t *value) {
    if (!ctx || !value) return jp_fail(JP_ERR_INVALID, "jp_get_option: null argument");
    if (option == JP_OPT_MOVE_MODE) { *value = ctx->move_mode; return JP_OK; }
    if (option == JP_OPT_P2G_MODE) { *value = ctx->p2g_mode; return JP_OK; }
    if (option == JP_OPT_ADVECT_AFFINE) { *value = ctx->g.affine; return JP_OK; }
    if (option == JP_OPT_MOVE_POLICY) { *value = ctx->move_policy; return JP_OK; }
    if (option == JP_OPT_GRAPH_STEP_OFFSET) {
        unsigned int v = 0;
        JP_CUDA(cudaSetDevice(ctx->device));
        JP_CUDA(cudaMemcpy(&v, jp_step_offset(ctx->stats), sizeof(v), cudaMemcpyDeviceToHost));
        *value = (int32_t)v;
        return JP_OK;
    }
    if (option == JP_OPT_ADVECT_CLASSIFY) { *value = ctx->hint_opt; return JP_OK; }
    if (option == JP_OPT_LAST_CLASSIFY) { *value = ctx->last_classify; return JP_OK; }
    if (option == JP_OPT_MOVE_INTERP) { *value = ctx->mi_opt; return JP_OK; }
    if (option == JP_OPT_LAST_INTERP) { *value = (ctx->last_p2g_handoff ? 1 : 0) | (ctx->last_phase_handoff ? 2 : 0); return JP_OK; }
    return jp_fail(JP_ERR_INVALID, "jp_get_option: unknown option");
}

static bool mi_matches(const jp_ctx *ctx, const jp_particles *p) {
    if (ctx->mi_key[3] != (const void *)p->index) return false;
    for (int d = 0; d < ctx->g.ndim; d++) if (ctx->mi_key[d] != (const void *)p->coords[d]) return false;
    return true;
}

extern "C" int jp_particle2grid(jp_ctx *ctx, const jp_particles *p, double *F, const double *Fp, void *stream) {
    PREP("jp_particle2grid");
    if (!Fp || !F) return jp_fail(JP_ERR_INVALID, "jp_particle2grid: null field");
    const dim3 ng = tile_grid(g.n[0] + 1, g.n[1] + 1, g.ndim == 3 ? g.n[2] + 1 : 1);
    if (ctx->p2g_mode == JP_P2G_EXACT || g.S > JP_MAX_SLOTS) {     // the two-pass cell kernel keeps a 64-bit mask per cell
        if (g.ndim == 2) k_p2g<2><<<ng, blk, 0, st>>>(g, cco, p->index, F, Fp);
        else             k_p2g<3><<<ng, blk, 0, st>>>(g, cco, p->index, F, Fp);
    } else {
        const int NQ = g.ndim == 2 ? 4 : 8;
        int rc = p2g_ws_reserve(ctx);
        if (rc) return rc;
        double *PW = ctx->p2g_ws, *PWF = ctx->p2g_ws + (int64_t)NQ * g.C;
        const bool fw = ctx->p2g_mode == JP_P2G_TWOPASS_FASTW;
        // move -> interpolation hand-off: the cell sums of this field were left by the last jp_move's scatter pass; the cell pass
        // then runs only if that move took the direct sweeps after all (device-side flag)
        const unsigned int *run_flag = nullptr;
        if (ctx->mi_valid_p2g && Fp == ctx->mi_fp && ctx->mi_p2g_mode == ctx->p2g_mode && mi_matches(ctx, p)) run_flag = ctx->mp_flag;
        else ctx->mi_valid_p2g = 0;                                 // the workspace is overwritten with another field's sums
        ctx->last_p2g_handoff = run_flag != nullptr;
        if (g.ndim == 2) {
            if (fw) k_p2g_cell<2, true><<<grd, blk, 0, st>>>(g, cco, p->index, Fp, PW, PWF, run_flag);
            else    k_p2g_cell<2, false><<<grd, blk, 0, st>>>(g, cco, p->index, Fp, PW, PWF, run_flag);
            k_p2g_node<2><<<ng, blk, 0, st>>>(g, PW, PWF, F);
        } else {
            if (fw) k_p2g_cell<3, true><<<grd, blk, 0, st>>>(g, cco, p->index, Fp, PW, PWF, run_flag);
            else    k_p2g_cell<3, false><<<grd, blk, 0, st>>>(g, cco, p->index, Fp, PW, PWF, run_flag);
            k_p2g_node<3><<<ng, blk, 0, st>>>(g, PW, PWF, F);
        }
    }
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_particle2centroid(jp_ctx *ctx, const jp_particles *p, double *Fc, const double *Fp, void *stream) {
    PREP("jp_particle2centroid");
    if (!Fp || !Fc) return jp_fail(JP_ERR_INVALID, "jp_particle2centroid: null field");
    if (g.ndim == 2) k_p2c<2><<<grd, blk, 0, st>>>(g, cco, Fc, Fp);
    else             k_p2c<3><<<grd, blk, 0, st>>>(g, cco, Fc, Fp);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

template <int N>
static void launch_phase(const JpGrid &g, dim3 grd, dim3 blk, cudaStream_t st, CPtr3 cco, double *ratios, const double *phases, int K,
                         const unsigned int *run_flag) {
    if (K <= 2) k_phase<N, 2><<<grd, blk, 0, st>>>(g, cco, ratios, phases, K, run_flag);
    else if (K <= 4) k_phase<N, 4><<<grd, blk, 0, st>>>(g, cco, ratios, phases, K, run_flag);
    else if (K <= 8) k_phase<N, 8><<<grd, blk, 0, st>>>(g, cco, ratios, phases, K, run_flag);
    else if (K <= 16) k_phase<N, 16><<<grd, blk, 0, st>>>(g, cco, ratios, phases, K, run_flag);
    else k_phase<N, 32><<<grd, blk, 0, st>>>(g, cco, ratios, phases, K, run_flag);
}

extern "C" int jp_phase_ratios_center(jp_ctx *ctx, const jp_particles *p, double *ratios, const double *phases, int32_t K, void *stream) {
    PREP("jp_phase_ratios_center");
    if (!ratios || !phases) return jp_fail(JP_ERR_INVALID, "jp_phase_ratios_center: null field");
    if (K < 1 || K > JP_MAX_PHASES) return jp_fail(JP_ERR_UNSUPPORTED, "jp_phase_ratios_center: 1 <= nphases <= 32 required");
    // move -> interpolation hand-off: the ratios were left by the last jp_move's scatter pass (k_phase runs only if that move
    // took the direct sweeps after all)
    const unsigned int *run_flag = nullptr;
    if (ctx->mi_valid_ph && phases == ctx->mi_ph && K == ctx->mi_K && mi_matches(ctx, p)) run_flag = ctx->mp_flag;
    ctx->last_phase_handoff = run_flag != nullptr;
    if (g.ndim == 2) launch_phase<2>(g, grd, blk, st, cco, ratios, phases, K, run_flag);
    else             launch_phase<3>(g, grd, blk, st, cco, ratios, phases, K, run_flag);
    if (run_flag) k_copy_unless<<<148 * 8, 256, 0, st>>>(ratios, ctx->mi_rc, (int64_t)K * g.C, run_flag);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

// K-dispatch shared by the vertex / face / midpoint launchers
#define JP_PHASE_DISPATCH(KERNEL, NTPL, GRID, ...)                                              \
    do {                                                                                        \
        if (K <= 2) KERNEL<NTPL 2><<<GRID, blk, 0, st>>>(__VA_ARGS__);                          \
        else if (K <= 4) KERNEL<NTPL 4><<<GRID, blk, 0, st>>>(__VA_ARGS__);                     \
        else if (K <= 8) KERNEL<NTPL 8><<<GRID, blk, 0, st>>>(__VA_ARGS__);                     \
        else if (K <= 16) KERNEL<NTPL 16><<<GRID, blk, 0, st>>>(__VA_ARGS__);                   \
        else KERNEL<NTPL 32><<<GRID, blk, 0, st>>>(__VA_ARGS__);                                \
    } while (0)
#define JP_COMMA ,

extern "C" int jp_phase_ratios_vertex(jp_ctx *ctx, const jp_particles *p, double *ratios, const double *phases, int32_t K, void *stream) {
    PREP("jp_phase_ratios_vertex");
    if (!ratios || !phases) return jp_fail(JP_ERR_INVALID, "jp_phase_ratios_vertex: null field");
    if (K < 1 || K > JP_MAX_PHASES) return jp_fail(JP_ERR_UNSUPPORTED, "jp_phase_ratios_vertex: 1 <= nphases <= 32 required");
    const dim3 ng = tile_grid(g.n[0] + 1, g.n[1] + 1, g.ndim == 3 ? g.n[2] + 1 : 1);
    if (g.ndim == 2) JP_PHASE_DISPATCH(k_phase_vertex, 2 JP_COMMA, ng, g, cco, ratios, phases, K);
    else             JP_PHASE_DISPATCH(k_phase_vertex, 3 JP_COMMA, ng, g, cco, ratios, phases, K);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_phase_ratios_face(jp_ctx *ctx, const jp_particles *p, double *ratios, const double *phases, int32_t K, int32_t dim, void *stream) {
    PREP("jp_phase_ratios_face");
    if (!ratios || !phases) return jp_fail(JP_ERR_INVALID, "jp_phase_ratios_face: null field");
    if (K < 1 || K > JP_MAX_PHASES) return jp_fail(JP_ERR_UNSUPPORTED, "jp_phase_ratios_face: 1 <= nphases <= 32 required");
    if (dim < 0 || dim >= g.ndim) return jp_fail(JP_ERR_INVALID, "jp_phase_ratios_face: dimension must be :x, :y or :z");
    if (g.ndim == 2) JP_PHASE_DISPATCH(k_phase_face, 2 JP_COMMA, grd, g, cco, ratios, phases, K, dim, false);
    else             JP_PHASE_DISPATCH(k_phase_face, 3 JP_COMMA, grd, g, cco, ratios, phases, K, dim, false);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_phase_ratios_midpoint(jp_ctx *ctx, const jp_particles *p, double *ratios, const double *phases, int32_t K, int32_t plane, void *stream) {
    PREP("jp_phase_ratios_midpoint");
    if (!ratios || !phases) return jp_fail(JP_ERR_INVALID, "jp_phase_ratios_midpoint: null field");
    if (K < 1 || K > JP_MAX_PHASES) return jp_fail(JP_ERR_UNSUPPORTED, "jp_phase_ratios_midpoint: 1 <= nphases <= 32 required");
    if (g.ndim != 3) return jp_fail(JP_ERR_INVALID, "jp_phase_ratios_midpoint: 3-D only");
    if (plane < 0 || plane > 2) return jp_fail(JP_ERR_INVALID, "jp_phase_ratios_midpoint: Unknown dimensions. Valid dimensions are :xy, :yz, :xz");
    JP_PHASE_DISPATCH(k_phase_midpoint, , grd, g, cco, ratios, phases, K, plane, -1, -1);
    JP_CHECK_LAUNCH();
    return JP_OK;
}

// update_phase_ratios!(phase_ratios, particles, phases) (src/PhaseRatios/utils.jl:15-41) in one call.
// mode JP_PHASE_LITERAL: the reference's sequence of kernels (bit-exact).  JP_PHASE_FUSED: one pass over the
// particles + node gathers (csrc/jp_phase_ratios.cuh), within the stated 1e-12; needs K <= 4, else literal.
template <int N, int KMAX>
static int update_phase_ratios_fused(jp_ctx *ctx, const jp_particles *p, const double *phases, int K, double *center, double *vertex,
                                     double *const *faces, double *const *mids, cudaStream_t st) {
    using P = PhaseFused<N>;
    const JpGrid &g = ctx->g;
    const size_t need = (size_t)P::NT * K * g.C;
    if (need > ctx->pr_ws_elems) {
        JP_NO_ALLOC_IN_CAPTURE("update_phase_ratios!: the per-cell weight workspace");
        if (ctx->pr_ws) JP_CUDA(cudaFree(ctx->pr_ws));
        ctx->pr_ws = nullptr; ctx->pr_ws_elems = 0;
        JP_CUDA(cudaMalloc(&ctx->pr_ws, need * sizeof(double)));
        ctx->pr_ws_elems = need;
    }
    CPtr3 cco = {{p->coords[0], p->coords[1], p->coords[2]}};
    const size_t smem = (size_t)P::NT * KMAX * P::THREADS * sizeof(double);
    static bool configured = false;
    if (!configured) {
        JP_CUDA(cudaFuncSetAttribute(k_phase_fused_cell<N, KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const dim3 cblk(32, 4, 1), cgrd((g.n[0] + 31) / 32, (g.n[1] + 3) / 4, N == 3 ? g.n[2] : 1);
    k_phase_fused_cell<N, KMAX><<<cgrd, cblk, smem, st>>>(g, cco, phases, K, center, ctx->pr_ws);
    JP_CHECK_LAUNCH();
    const dim3 blk(JP_BX, JP_BY, 1);
    k_phase_fused_node<N, KMAX><<<tile_grid(g.n[0] + 1, g.n[1] + 1, N == 3 ? g.n[2] + 1 : 1), blk, 0, st>>>(g, ctx->pr_ws, vertex, K, 0, 0);
    const dim3 grd = tile_grid(g.n[0], g.n[1], g.n[2]);
    for (int d = 0; d < N; d++) {
        const int nn[3] = {g.n[0] + (d == 0), g.n[1] + (d == 1), N == 3 ? g.n[2] + (d == 2) : 1};
        k_phase_fused_node<N, KMAX><<<tile_grid(nn[0], nn[1], nn[2]), blk, 0, st>>>(g, ctx->pr_ws, faces[d], K, 1, d);
        k_phase_face<N, KMAX><<<(unsigned)((g.C / g.n[d] + 255) / 256), blk, 0, st>>>(g, cco, faces[d], phases, K, d, true);   // low-boundary faces: literal branch
    }
    if (N == 3)
        for (int pl = 0; pl < 3; pl++) {
            const int off[3] = {pl != 1, pl != 2, pl != 0};
            k_phase_fused_node<N, KMAX><<<tile_grid(g.n[0] + off[0], g.n[1] + off[1], g.n[2] + off[2]), blk, 0, st>>>(g, ctx->pr_ws, mids[pl], K, 2, pl);
            const int da = pl == 1 ? 1 : 0, db = pl == 0 ? 1 : 2;                                  // low-boundary midpoints: literal branch on the two planes
            k_phase_midpoint<KMAX><<<(unsigned)((g.C / g.n[da] + 255) / 256), blk, 0, st>>>(g, cco, mids[pl], phases, K, pl, da, -1);
            k_phase_midpoint<KMAX><<<(unsigned)((g.C / g.n[db] + 255) / 256), blk, 0, st>>>(g, cco, mids[pl], phases, K, pl, db, da);
        }
    JP_CHECK_LAUNCH();
    return JP_OK;
}

extern "C" int jp_update_phase_ratios(jp_ctx *ctx, const jp_particles *p, const double *phases, int32_t K, double *center, double *vertex,
                                      double *const *faces, double *const *midpoints, int32_t mode, void *stream) {
    PREP("jp_update_phase_ratios");
    if (!phases || !center || !vertex || !faces) return jp_fail(JP_ERR_INVALID, "jp_update_phase_ratios: null field");
    for (int d = 0; d < g.ndim; d++) if (!faces[d]) return jp_fail(JP_ERR_INVALID, "jp_update_phase_ratios: null face field");
    if (g.ndim == 3) { if (!midpoints) return jp_fail(JP_ERR_INVALID, "jp_update_phase_ratios: null midpoint fields");
                       for (int d = 0; d < 3; d++) if (!midpoints[d]) return jp_fail(JP_ERR_INVALID, "jp_update_phase_ratios: null midpoint field"); }
    if (K < 1 || K > JP_MAX_PHASES) return jp_fail(JP_ERR_UNSUPPORTED, "jp_update_phase_ratios: 1 <= nphases <= 32 required");
    if (mode != JP_PHASE_LITERAL && mode != JP_PHASE_FUSED) return jp_fail(JP_ERR_INVALID, "jp_update_phase_ratios: unknown mode");
    if (mode == JP_PHASE_FUSED && K <= 4) {
        if (g.ndim == 2) return K <= 2 ? update_phase_ratios_fused<2, 2>(ctx, p, phases, K, center, vertex, faces, midpoints, st)
                                       : update_phase_ratios_fused<2, 4>(ctx, p, phases, K, center, vertex, faces, midpoints, st);
        return K <= 2 ? update_phase_ratios_fused<3, 2>(ctx, p, phases, K, center, vertex, faces, midpoints, st)
                      : update_phase_ratios_fused<3, 4>(ctx, p, phases, K, center, vertex, faces, midpoints, st);
    }
    int rc = jp_phase_ratios_center(ctx, p, center, phases, K, stream);
    if (!rc) rc = jp_phase_ratios_vertex(ctx, p, vertex, phases, K, stream);
    for (int d = 0; d < g.ndim && !rc; d++) rc = jp_phase_ratios_face(ctx, p, faces[d], phases, K, d, stream);
    if (g.ndim == 3) for (int pl = 0; pl < 3 && !rc; pl++) rc = jp_phase_ratios_midpoint(ctx, p, midpoints[pl], phases, K, pl, stream);
    return rc;
}

static int64_t plane_cells(const JpGrid &g, int dim) {
    return dim == 0 ? (int64_t)g.n[1] * g.n[2] : dim == 1 ? (int64_t)g.n[0] * g.n[2] : (int64_t)g.n[0] * g.n[1];
}
extern "C" int64_t jp_halo_plane_bytes(const jp_ctx *ctx, int32_t dim, int32_t narrays) {
    if (!ctx || dim < 0 || dim >= ctx->g.ndim || narrays < 0) return -1;
    return plane_cells(ctx->g, dim) * ctx->g.S * (8 * (int64_t)narrays + 1);
}
// a plane was rewritten after the advection -> move hand-off was left: its classification words are stale
static void halo_mark_dirty(jp_ctx *ctx, int dim, int plane) {
    ctx->bucketed = 0;                                       // a plane of the neighbour's advected particles
    if (!(ctx->hint_valid || (ctx->adv_split && ctx->hint_opt))) return;
    for (int i = 0; i < ctx->hint_ndirty; i++)
        if (ctx->hint_dirty[i][0] == dim && ctx->hint_dirty[i][1] == plane) return;
    if (ctx->hint_ndirty < 8) { ctx->hint_dirty[ctx->hint_ndirty][0] = dim; ctx->hint_dirty[ctx->hint_ndirty][1] = plane; ctx->hint_ndirty++; }
    else { hint_invalidate(ctx); if (ctx->adv_split) ctx->adv_split = 2; }     // 2: this step's hand-off is void
}

#include "jp_halo_nccl.cuh"

static int halo_common(jp_ctx *ctx, int dim, int plane, double *const *arrays, int narrays, uint8_t *index, void *buf, void *stream, bool pack) {
    if (!ctx || !buf || !index) return jp_fail(JP_ERR_INVALID, "jp_halo: null argument");
    const JpGrid &g = ctx->g;
    if (dim < 0 || dim >= g.ndim || plane < 0 || plane >= g.n[dim]) return jp_fail(JP_ERR_INVALID, "jp_halo: bad dim/plane");
    if (narrays < 0 || narrays > JP_MAX_ARGS + 3) return jp_fail(JP_ERR_UNSUPPORTED, "jp_halo: too many arrays");
    HaloArrs h; h.n = narrays;
    for (int a = 0; a < narrays; a++) {
        if (!arrays || !arrays[a]) return jp_fail(JP_ERR_INVALID, "jp_halo: null array");
        h.a[a] = arrays[a];
    }
    JP_CUDA(cudaSetDevice(ctx->device));
    const int64_t M = plane_cells(g, dim);
    const dim3 grd((unsigned)((M + 255) / 256 < 64 ? (M + 255) / 256 : 64), (unsigned)((narrays + 1) * g.S));
    if (pack) k_halo_plane<true><<<grd, 256, 0, (cudaStream_t)stream>>>(g, dim, plane, h, index, (unsigned char *)buf, (int)M);
    else {
        mi_invalidate(ctx);
        k_halo_plane<false><<<grd, 256, 0, (cudaStream_t)stream>>>(g, dim, plane, h, index, (unsigned char *)buf, (int)M);
        halo_mark_dirty(ctx, dim, plane);
    }
    JP_CHECK_LAUNCH();
    return JP_OK;
}
extern "C" int jp_halo_pack(jp_ctx *ctx, int32_t dim, int32_t plane, double *const *arrays, int32_t narrays, const uint8_t *index, void *buf, void *stream) {
    return halo_common(ctx, dim, plane, arrays, narrays, (uint8_t *)index, buf, stream, true);
}
extern "C" int jp_halo_unpack(jp_ctx *ctx, int32_t dim, int32_t plane, double *const *arrays, int32_t narrays, uint8_t *index, const void *buf, void *stream) {
    return halo_common(ctx, dim, plane, arrays, narrays, index, (void *)buf, stream, false);
}

// ---------------------------------------------------------------------------
// Array(CellArray) / CuArray(CellArray) layout conversion (jp_convert.cuh)
extern "C" int jp_cellarray_permute(jp_ctx *ctx, const void *src, int32_t src_type, void *dst, int32_t dst_type, int64_t ncells,
                                    int32_t ncomp, int32_t direction, void *stream) {
    if (!src || !dst) return jp_fail(JP_ERR_INVALID, "jp_cellarray_permute: null argument");
    if (src == dst) return jp_fail(JP_ERR_INVALID, "jp_cellarray_permute: in-place conversion is not supported");
    if (ncells < 0 || ncomp < 0) return jp_fail(JP_ERR_INVALID, "jp_cellarray_permute: negative extent");
    if (direction != JP_LAYOUT_TO_HOST && direction != JP_LAYOUT_TO_DEVICE) return jp_fail(JP_ERR_INVALID, "jp_cellarray_permute: unknown direction");
    const bool sb = src_type == JP_BOOL, db = dst_type == JP_BOOL;
    if (src_type < JP_F64 || src_type > JP_BOOL || dst_type < JP_F64 || dst_type > JP_BOOL || sb != db)
        return jp_fail(JP_ERR_UNSUPPORTED, "jp_cellarray_permute: element types must be Float64/Float32 (convertible) or Bool -> Bool");
    if (ncells == 0 || ncomp == 0) return JP_OK;
    if ((ncells + 31) / 32 > 2147483647LL) return jp_fail(JP_ERR_UNSUPPORTED, "jp_cellarray_permute: too many cells");
    if (ctx) JP_CUDA(cudaSetDevice(ctx->device));       // no context: the caller's current device
    cudaStream_t st = (cudaStream_t)stream;
    const int th = direction == JP_LAYOUT_TO_HOST;
    if (sb) launch_cellarray_permute<uint8_t, uint8_t>(src, dst, ncells, ncomp, th, st);
    else if (src_type == JP_F64 && dst_type == JP_F64) launch_cellarray_permute<double, double>(src, dst, ncells, ncomp, th, st);
    else if (src_type == JP_F64 && dst_type == JP_F32) launch_cellarray_permute<double, float>(src, dst, ncells, ncomp, th, st);
    else if (src_type == JP_F32 && dst_type == JP_F64) launch_cellarray_permute<float, double>(src, dst, ncells, ncomp, th, st);
    else launch_cellarray_permute<float, float>(src, dst, ncells, ncomp, th, st);
    JP_CHECK_LAUNCH();
    return JP_OK;
}
