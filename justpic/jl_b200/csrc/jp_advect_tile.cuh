// jp_advect_tile.cuh -- advection! for the standard staggered layout
// (xvel[c][d] = vertex vector if c == d, ghosted-centre vector otherwise).
//
// One CTA owns a brick of TX x TY x TZ cells (one warp per x-run of 32 cells):
//  1. the brick's velocity stencils -- origin one node below the brick, extent
//     T+4 nodes per dimension, which covers every stage position of a particle
//     displaced by at most one cell -- are staged in shared memory together with
//     the grid-vector segments, so the 2^N-corner gathers of every RK stage are
//     LDS with compile-time strides instead of 64-bit-addressed global loads;
//  2. each warp ballots its cells' occupancy masks slot by slot into a small
//     shared-memory ring of compacted (slot, cell) entries, consumed 32 entries at
//     a time, so every lane carries a live particle even when the slot
//     planes are half empty (the reference's first-free-slot policy leaves them at
//     ~50-60 % occupancy), while loads stay coalesced because the list is
//     slot-major;
//  3. a particle whose stage position leaves the staged stencil, sits exactly on
//     a grid node, or is NaN/outside the domain takes the literal global-memory
//     routine (jp_interp_velocity_literal), so results are bitwise those of the
//     literal code in every case.
#pragma once
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include "jp_core.h"

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers --------------------------------------
struct AdvTmaMaps { CUtensorMap m[3]; };

__device__ __forceinline__ void adv_mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void adv_mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void adv_mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_LOOP;\n\t}"
        ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void adv_tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)),
                   "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void adv_tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"((unsigned)__cvta_generic_to_shared(bar)),
                   "r"(x), "r"(y) : "memory");
}

// H = 1: one more staged node below and two more above in y / z, for the LinP / MQS stencils (interpolation cell -1 .. +2)
template <int N, int H = 0> struct AdvTile {
    static constexpr int TX = 32;
#ifndef JP_ADV_TPSM
#define JP_ADV_TPSM 512      // resident threads per SM the register budget is sized for (r02: 2 CTAs of 256 at <= 128 registers beat 3 at 80)
#endif
#ifndef JP_ADV_TY
#define JP_ADV_TY 4
#define JP_ADV_TZ 2
#define JP_ADV_EX 38
#endif
    static constexpr int TY = N == 3 ? JP_ADV_TY : 8;
    static constexpr int TZ = N == 3 ? JP_ADV_TZ : 1;
    // staged nodes per dim: one node below the brick in y/z; TWO below in x so that the box starts
    // on an even node -- TMA needs the box origin 16-byte aligned in the innermost dimension.
    // The x extent (needed: TX + 5) is padded to 48 so that the row pitch is a multiple of 16 doubles:
    // every stencil row then starts on bank 0 and a lane's bank depends on its x index only.
    static constexpr int OX = 2;
    static constexpr int EX = N == 3 ? JP_ADV_EX : 40, EY = TY + 4 + 2 * H, EZ = N == 3 ? TZ + 4 + 2 * H : 1;
    static constexpr int VOL = ((EX * EY * EZ + 15) / 16) * 16;                 // tile pitch: multiple of 128 bytes
    static constexpr int NW = TY * TZ;                                          // warps per CTA
};

// shared-memory layout (doubles first, then the uint16 work lists)
template <int N, bool UNIFORM, int H = 0> struct AdvSmem {
    using T = AdvTile<N, H>;
    static constexpr int V_OFF = 0;                       // N tiles of VOL doubles
    static constexpr int VEC = 40;                        // grid-vector segment per dim (needed: TX + 6)
    static constexpr int XV_OFF = N * T::VOL;
    static constexpr int XG_OFF = XV_OFF + 3 * VEC;
    static constexpr int IXV_OFF = XG_OFF + 3 * VEC;      // reciprocal spacings: non-uniform grids only
    static constexpr int IXG_OFF = IXV_OFF + 3 * VEC;
    static constexpr int NDOUBLES = UNIFORM ? IXV_OFF : IXG_OFF + 3 * VEC;
    static constexpr int RING = 128;                      // per-warp ring of compacted (slot, cell) entries
    static constexpr size_t BYTES = sizeof(double) * NDOUBLES + sizeof(uint16_t) * T::NW * RING;
};

// velocity at p from the staged stencils; r0[d] = tile-relative index of the seed
// (storage) cell, c0[d] = global index of the tile's first staged node.
// AFFINE (implies UNIFORM; 1: vertex vectors, 2: vertex and ghosted-centre vectors): grid coordinates
// are regenerated as fma(i, dx, x0) -- verified on the host to reproduce every stored entry bit for
// bit -- instead of being loaded from shared memory:
// the kernel is bound by LDS return bandwidth (128 B/clk/SM) and 30 of its 78 loads per particle
// were grid-vector entries.
// Stage 1: the first interpolation of a particle (at its own position) skips the re-centring block when the library knows
// that every particle lies strictly inside its storage cell (g.bucketed: the last call that touched the particles was
// move_particles!, init, inject or clean -- not another advection!, a halo unpack or a foreign write); a particle that is not is
// flagged and takes the literal routine (same result).  Read at run time the flag costs 0.29 ms at 256^3 (the re-centring code stays
// in the loop: profiles/r02ad_ab_bucketed_flag.log), so the variants the time loops use -- trilinear, range grids -- exist twice:
// BKT = true is launched when the flag is set and has the shortcut compiled in; all others test g.bucketed.
// Later stages: they run the re-centring arithmetic unconditionally -- at CFL 0.5 nearly every warp holds
// a lane whose stage position left its seed cell, so the vote only cost a divergent-branch frame.  Measured at 256^3 (r02l, all
// bit-identical): 14.44 -> 13.44 (stage 1) -> 13.32 (no vote) -> 13.09 ms (2 CTAs / SM at <= 128 registers instead of 3 at 80).
// Dropped: stencil loads as ld.shared.f64 on a 32-bit address held in one opaque register (the compiler re-derives the shared
// window base at every group of loads: S2R CgaCtaId, MOV, VIADD, LEA) -- 13.09 -> 13.32 ms; bricks of 32 x 8 x 2 / 32 x 4 x 4 cells
// (16 warps): 15.0 / 14.4 ms; 4 CTAs / SM at 64 registers: 17.5 ms; an extra stencil row per z-plane so that consecutive planes fall
// into different banks (the plane pitch is 0 mod 32 banks): 13.08 -> 13.32 ms; a second, one-node-shifted copy of every tile so that
// the two x-adjacent corners are one 16-byte LDS.128 (24 instead of 48 stencil loads per particle): 13.07 -> 13.67 ms.
// stencil values of the LinP / MQS interpolants from the staged tile: A(i1, j1, k1) with 1-based GLOBAL node indices (jp_core.h)
template <int N, class T> struct AdvTileAcc {
    const double *tile; int c0x, c0y, c0z;
    __device__ __forceinline__ double operator()(int i1, int j1, int k1) const {
        return tile[(i1 - 1 - c0x) + T::EX * ((j1 - 1 - c0y) + (N == 3 ? T::EY * (k1 - 1 - c0z) : 0))];
    }
};
template <int N, bool UNIFORM, int AFFINE, bool FIRST = false, int INTERP = 0, bool BKT = false>
__device__ __forceinline__ bool adv_interp_tile(const JpGrid &g, const double *__restrict__ sm, const int *c0, const int *r0,
                                                const double *gd0, const double *p, double *vout, unsigned amask) {
    using T = AdvTile<N, INTERP ? 1 : 0>;
    using L = AdvSmem<N, UNIFORM, INTERP ? 1 : 0>;
    // Straight-line code: a lane that cannot be served from the tile (tie with a grid node, NaN, more
    // than one cell from its seed, outside the domain) only raises `bad` and is redone by the literal
    // routine afterwards.  Divergent early exits made the compiler run the rest of the stage once per
    // path (stage 2 was issued twice per batch at ~16 active lanes).
    bool bad = false;
    int iv[3], ig[3];
    double tv[3], tg[3];
    double xav[3], xag[3], dxv[3], dxg[3];        // INTERP > 0: lower coordinate / spacing of the interpolation cell per grid kind
#pragma unroll
    for (int d = 0; d < N; d++) {
        const double *xv = sm + L::XV_OFF + d * L::VEC;
        const double *xg = sm + L::XG_OFF + d * L::VEC;
        const double pd = p[d];
        int r = r0[d];
        double gd = AFFINE ? gd0[d] : 0.0;                    // (double)(global index of node r)
        double a, b;
        if (AFFINE) { a = fma(gd, g.aff_dv[d], g.aff_v0[d]); b = fma(gd + 1.0, g.aff_dv[d], g.aff_v0[d]); }
        else { a = xv[r]; b = xv[r + 1]; }
        const bool up = pd > b, dn = pd < a;
        if (!(FIRST && (BKT || g.bucketed))) {    // first interpolation of a bucketed particle: nothing to re-centre; later stages: always (no vote)
            r += (up ? 1 : 0) - (dn ? 1 : 0);
            if (AFFINE) {
                gd += (up ? 1.0 : 0.0) - (dn ? 1.0 : 0.0);
                a = fma(gd, g.aff_dv[d], g.aff_v0[d]); b = fma(gd + 1.0, g.aff_dv[d], g.aff_v0[d]);
            } else { a = xv[r]; b = xv[r + 1]; }
            const int gi = c0[d] + r;
            bad = bad || gi < 0 || gi >= g.n[d];
        }
        bad = bad || !(a < pd && pd < b);                     // also catches ties and NaN
        iv[d] = r;
        tv[d] = (pd - a) * (UNIFORM ? g.inv_dv[d] : sm[L::IXV_OFF + d * L::VEC + r]);
        if (INTERP) { xav[d] = a; dxv[d] = UNIFORM ? g.dxv0[d] : b - a; }
        const double m = AFFINE == 2 ? fma(gd + 1.0, g.aff_dg[d], g.aff_g0[d]) : xg[r + 1];
        bad = bad || pd == m;
        const bool lower = pd < m;
        ig[d] = lower ? r : r + 1;
        const double gl = AFFINE == 2 ? fma(gd, g.aff_dg[d], g.aff_g0[d]) : xg[r];
        tg[d] = (pd - (lower ? gl : m)) * (UNIFORM ? g.inv_dg[d] : sm[L::IXG_OFF + d * L::VEC + ig[d]]);
        if (INTERP) { xag[d] = lower ? gl : m; dxg[d] = UNIFORM ? g.dxg0[d] : (lower ? m - gl : xg[r + 2] - m); }
    }
#pragma unroll
    for (int c = 0; c < N; c++) {
        const int ix = c == 0 ? iv[0] : ig[0];
        const int iy = c == 1 ? iv[1] : ig[1];
        const int iz = N == 3 ? (c == 2 ? iv[2] : ig[2]) : 0;
        const double t[3] = {c == 0 ? tv[0] : tg[0], c == 1 ? tv[1] : tg[1], N == 3 ? (c == 2 ? tv[2] : tg[2]) : 0.0};
        double v[8];
        const double *F = sm + L::V_OFF + c * T::VOL + ix + T::EX * (iy + T::EY * iz);
        v[0] = F[0]; v[1] = F[1]; v[2] = F[T::EX]; v[3] = F[T::EX + 1];
        if (N == 3) {
            v[4] = F[T::EX * T::EY]; v[5] = F[T::EX * T::EY + 1];
            v[6] = F[T::EX * T::EY + T::EX]; v[7] = F[T::EX * T::EY + T::EX + 1];
        }
        const double VL = jp_lerp<N>(v, t);
        if (INTERP == 0) vout[c] = VL;
        else {
            // advection_LinP! / advection_MQS!: the linear value plus a correction when the interpolation cell is interior
            // (1 < idx < size(F) - 1 in every direction, jp_interp_velocity_hi); the stencil idx - 1 .. idx + 2 must lie in the tile
            const int ti[3] = {ix, iy, iz};
            const int ext[3] = {T::EX, T::EY, T::EZ};
            int idx1[3] = {1, 1, 1};
            bool interior = true, intile = true;
#pragma unroll
            for (int d = 0; d < N; d++) {
                idx1[d] = c0[d] + ti[d] + 1;
                interior = interior && 1 < idx1[d] && idx1[d] < g.nvel[c][d] - 1;
                intile = intile && ti[d] >= 1 && ti[d] + 2 < ext[d];
            }
            vout[c] = VL;
            if (interior && !intile) bad = true;
            else if (interior) {
                const AdvTileAcc<N, T> acc = {sm + L::V_OFF + c * T::VOL, c0[0], c0[1], N == 3 ? c0[2] : 0};
                if (INTERP == 2) vout[c] = jp_mqs<N>(acc, g.nvel[c], c, idx1, v, t);
                else {
                    const double xcn[3] = {c == 0 ? xav[0] : xag[0], c == 1 ? xav[1] : xag[1], N == 3 ? (c == 2 ? xav[2] : xag[2]) : 0.0};
                    const double dxi[3] = {c == 0 ? dxv[0] : dxg[0], c == 1 ? dxv[1] : dxg[1], N == 3 ? (c == 2 ? dxv[2] : dxg[2]) : 1.0};
                    vout[c] = jp_linp<N>(acc, g.nvel[c], c, idx1, xcn, dxi, p, VL);
                }
            }
        }
    }
    return !bad;
}

template <int N, bool UNIFORM, int AFFINE, bool FIRST = false, int INTERP = 0, bool BKT = false>
__device__ __forceinline__ void adv_interp(const JpGrid &g, const double *__restrict__ sm, const double *const *V, const int *c0,
                                           const int *r0, const double *gd0, const int *cell1, const double *p, double *vout, unsigned amask) {
    if (adv_interp_tile<N, UNIFORM, AFFINE, FIRST, INTERP, BKT>(g, sm, c0, r0, gd0, p, vout, amask)) return;
    if (INTERP == 0) jp_interp_velocity_literal<N>(g, V, p, cell1, vout);
    else jp_interp_velocity_hi<N, INTERP>(g, V, p, cell1, vout);
}

// HINT (JP_OPT_ADVECT_CLASSIFY, the advection -> move hand-off): every new position is also classified
// for the following move_particles! -- jp_classify_fast / jp_classify_particle on the value being stored,
// i.e. exactly what k_move_classify3 would compute from memory -- into a per-warp byte table in shared
// memory ([cell of the x-run][slot], rows of ADV_HINT_ROW bytes preset to "stays").  When the warp has finished its
// x-run, lane = cell reads its row as 64-bit words, finds the leavers of each word with byte-parallel bit tricks
// (~12 instructions per 8 slots) and appends their codes -- the leavers only, ~9 of 48 slots -- to the packed code
// words of the move plan, in slot order, exactly as k_move_classify3 packs them; then the occupancy / leave words.
// jp_move then starts at the plan kernels: no pass over the coordinates, no intermediate plane in HBM.
#define ADV_HINT_ROW 72          // bytes per cell row: JP_MAX_SLOTS + 8 (18 words: 64-bit row loads of a half-warp hit 32 distinct banks)
template <int N, int SCHEME, bool UNIFORM, int AFFINE, bool HINT, int INTERP = 0, bool BKT = false>
__global__ void __launch_bounds__(AdvTile<N>::NW * 32, (UNIFORM ? JP_ADV_TPSM : 768) / (AdvTile<N>::NW * 32)) k_advect_tile(JpGrid g, Ptr3 co, const uint8_t *__restrict__ index, CPtr3 V,
                                                                     double alpha, double dt,
                                                                     const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                                                                     const __grid_constant__ CUtensorMap tm2, int tma_mask, MovePlanWs ws,
                                                                     unsigned int *complex_flag) {
    constexpr int H = INTERP ? 1 : 0;
    using T = AdvTile<N, H>;
    using L = AdvSmem<N, UNIFORM, H>;
    extern __shared__ __align__(128) unsigned char smem_raw[];   // TMA destinations: 128-byte aligned (VOL*8 is a multiple of 128)
    double *sm = reinterpret_cast<double *>(smem_raw);
    uint16_t *wl_all = reinterpret_cast<uint16_t *>(smem_raw + sizeof(double) * L::NDOUBLES);
    uint8_t *code_all = smem_raw + L::BYTES;                     // HINT: [warp][slot][32] classification bytes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // brick origin (cells) and first staged node (one below)
    const int b0[3] = {(int)blockIdx.x * T::TX, (int)blockIdx.y * T::TY, N == 3 ? (int)blockIdx.z * T::TZ : 0};
    const int c0[3] = {b0[0] - T::OX, b0[1] - 1 - H, N == 3 ? b0[2] - 1 - H : 0};
    if (g.region) {                         // CTA-uniform: shell bricks first, interior bricks while the halo planes travel
        const int ext[3] = {T::TX, T::TY, T::TZ};
        bool shell = false;
#pragma unroll
        for (int d = 0; d < N; d++) {
            const int hi = min(b0[d] + ext[d], g.n[d]);           // brick covers cells [b0, hi)
            shell = shell || b0[d] <= 1 || hi >= g.n[d] - 1;      // holds a cell of layers {0, 1, n-2, n-1}
        }
        if ((g.region == 1) != shell) return;
    }

    // ---- 1. stage velocity stencils and grid-vector segments.
    // Components whose array meets the TMA constraints (16-byte aligned base and row pitch) are
    // fetched as ONE 3-D (2-D) box per component by cp.async.bulk.tensor, out-of-range nodes
    // zero-filled by the hardware, completion signalled on an mbarrier; the others (e.g. Vx, whose
    // leading extent n+1 is odd) by cooperative coalesced loads.
    __shared__ __align__(8) uint64_t tma_bar;
    if (tma_mask) {
        if (tid == 0) adv_mbar_init(&tma_bar, 1);
        __syncthreads();
        if (tid == 0) {
            adv_mbar_expect_tx(&tma_bar, (unsigned)(__popc(tma_mask) * T::EX * T::EY * T::EZ * sizeof(double)));
            // (descriptors are addressed statically: they must stay in kernel-parameter space)
            if (N == 3) {
                if (tma_mask & 1) adv_tma_load_3d(sm + L::V_OFF + 0 * T::VOL, &tm0, &tma_bar, c0[0], c0[1], c0[2]);
                if (tma_mask & 2) adv_tma_load_3d(sm + L::V_OFF + 1 * T::VOL, &tm1, &tma_bar, c0[0], c0[1], c0[2]);
                if (tma_mask & 4) adv_tma_load_3d(sm + L::V_OFF + 2 * T::VOL, &tm2, &tma_bar, c0[0], c0[1], c0[2]);
            } else {
                if (tma_mask & 1) adv_tma_load_2d(sm + L::V_OFF + 0 * T::VOL, &tm0, &tma_bar, c0[0], c0[1]);
                if (tma_mask & 2) adv_tma_load_2d(sm + L::V_OFF + 1 * T::VOL, &tm1, &tma_bar, c0[0], c0[1]);
            }
        }
    }
    for (int c = 0; c < N; c++) {
        if ((tma_mask >> c) & 1) continue;
        const double *__restrict__ F = V.p[c];
        const int n0 = g.nvel[c][0], n1 = g.nvel[c][1], n2 = N == 3 ? g.nvel[c][2] : 1;
        for (int t = tid; t < T::EX * T::EY * T::EZ; t += T::NW * 32) {
            const int ix = t % T::EX, iy = (t / T::EX) % T::EY, iz = t / (T::EX * T::EY);
            const int gx = c0[0] + ix, gy = c0[1] + iy, gz = N == 3 ? c0[2] + iz : 0;
            double val = 0.0;
            if (gx >= 0 && gx < n0 && gy >= 0 && gy < n1 && gz >= 0 && gz < n2)
                val = F[gx + (int64_t)n0 * (gy + (int64_t)n1 * gz)];
            sm[L::V_OFF + c * T::VOL + t] = val;
        }
    }
    if (HINT) {                                                   // this warp's rows of the byte table: every slot "stays" until a lane says otherwise
        uint64_t *rows = reinterpret_cast<uint64_t *>(code_all + (size_t)warp * 32 * ADV_HINT_ROW);
        for (int t = lane; t < 32 * ADV_HINT_ROW / 8; t += 32) rows[t] = 0x0101010101010101ull * JP_CLS_STAY;
    }
    if (tid < 3 * L::VEC) {
        const int d = tid / L::VEC, j = tid % L::VEC;
        if (d < N) {
            const int gi = c0[d] + j;
            sm[L::XV_OFF + tid] = (gi >= 0 && gi <= g.n[d]) ? g.xv[d][gi] : NAN;
            sm[L::XG_OFF + tid] = (gi >= 0 && gi <= g.n[d] + 1) ? g.xg[d][gi] : NAN;
            if (!UNIFORM) {
                sm[L::IXV_OFF + tid] = (gi >= 0 && gi < g.n[d]) ? g.ixv[d][gi] : NAN;
                sm[L::IXG_OFF + tid] = (gi >= 0 && gi <= g.n[d]) ? g.ixg[d][gi] : NAN;
            }
        }
    }

    // ---- 2. this warp's x-run of cells: occupancy masks -> compacted work list
    const int wy = warp % T::TY, wz = warp / T::TY;
    const int cy = b0[1] + wy, cz = b0[2] + wz;
    const int cx = b0[0] + lane;
    const bool row_ok = cy < g.n[1] && (N == 2 || cz < g.n[2]);
    const bool ok = row_ok && cx < g.n[0];
    const int64_t crow = (int64_t)g.n[0] * (cy + (N == 3 ? (int64_t)g.n[1] * cz : 0));   // linear index of cell (0, cy, cz)
    uint64_t m = 0;
    if (ok) {
        const uint8_t *ip = index + crow + cx;
#pragma unroll 8
        for (int s = 0; s < g.S; s++) m |= (uint64_t)(ip[(int64_t)s * g.C] != 0) << s;
    }
    if (tma_mask) adv_mbar_wait(&tma_bar, 0);   // TMA boxes have landed
    __syncthreads();                      // stencils staged by all warps

    // ---- 3. compaction + integration.  Slot planes are ~50 % full, so live (slot, cell) entries are
    //         ballot-compacted into a small shared-memory ring and integrated 32 at a time.  The two
    //         HALF-warps compact independently (lanes 0-15 <-> cells 0-15 of the x-run, 16-31 <-> 16-31):
    //         a 64-bit LDS is served per half-warp, and with all 16 lanes of a half inside a 16-cell
    //         window (stencil rows are bank-aligned, EX = 0 mod 16) their words fall in distinct banks;
    //         compacting across the whole warp spreads a half over up to 32 cells and two lanes 16 cells
    //         apart collide on every stencil load.
    constexpr int RH = L::RING / 2;
    // this half-warp's ring, addressed through a 32-bit shared-window address (one register, no generic->shared math per access)
    const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(wl_all + warp * L::RING + (lane >> 4) * RH);
    const unsigned lth = ((1u << (lane & 15)) - 1u) << (lane & 16);             // lower lanes of my half
    const double *Vp[3] = {V.p[0], V.p[1], V.p[2]};
    const bool hi = lane >> 4;
    int head0 = 0, tail0 = 0, head1 = 0, tail1 = 0, q = 0;
    // software pipeline: the coordinates of batch b+1 are fetched before batch b is integrated, so the
    // global-load latency (27 % of the stall samples before) overlaps the ~450 instructions of a batch
    bool cur_valid = false;
    int cur_l = 0;                     // ring entry: (slot << 5) | cell of the x-run
    int64_t cur_e = 0;
    double cur_p[3] = {0.0, 0.0, 0.0};
    for (;;) {
        // produce until both halves hold a full batch (or one ring is about to fill up)
        while (q < g.S && (tail0 - head0 < 16 || tail1 - head1 < 16) && tail0 - head0 <= RH - 16 && tail1 - head1 <= RH - 16) {
            const bool live = (m >> q) & 1ull;
            const unsigned bal = __ballot_sync(0xffffffffu, live);
            if (live) {
                const unsigned pos = ((hi ? tail1 : tail0) + __popc(bal & lth)) & (RH - 1);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(ring_sa + 2u * pos), "h"((unsigned short)((q << 5) | lane)) : "memory");
            }
            tail0 += __popc(bal & 0xffffu);
            tail1 += __popc(bal >> 16);
            q++;
        }
        __syncwarp();
        // fetch the next batch: ring entry -> element index -> coordinate loads (consumed one iteration later)
        const int k = (hi ? head1 : head0) + (lane & 15);
        const bool nxt_valid = k < (hi ? tail1 : tail0);
        int nxt_l = 0;
        int64_t nxt_e = 0;
        double nxt_p[3] = {0.0, 0.0, 0.0};
        if (nxt_valid) {
            unsigned short ent16;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(ent16) : "r"(ring_sa + 2u * (unsigned)(k & (RH - 1))) : "memory");
            nxt_l = ent16;
            nxt_e = crow + b0[0] + (ent16 & 31) + (int64_t)(ent16 >> 5) * g.C;
#pragma unroll
            for (int d = 0; d < N; d++) nxt_p[d] = co.p[d][nxt_e];
        }
        head0 = min(head0 + 16, tail0);
        head1 = min(head1 + 16, tail1);
        const bool any_next = __ballot_sync(0xffffffffu, nxt_valid) != 0;
        {
            const unsigned amask = __ballot_sync(0xffffffffu, cur_valid);
            if (cur_valid) {
                const int l = cur_l & 31;
                const int64_t e = cur_e;
                const int r0[3] = {l + T::OX, wy + 1 + H, wz + 1 + H};
                const int cell1[3] = {b0[0] + l + 1, cy + 1, cz + 1};
                const double gd0[3] = {(double)(b0[0] + l), (double)cy, (double)cz};     // global index of the seed cell's lower node
                double p0[3], k1[3], k2[3], qq[3], pn[3];
#pragma unroll
                for (int d = 0; d < N; d++) p0[d] = cur_p[d];
                adv_interp<N, UNIFORM, AFFINE, true, INTERP, BKT>(g, sm, Vp, c0, r0, gd0, cell1, p0, k1, amask);
                if (SCHEME == 0) {
                    const double cdt = 1.0 * dt;
#pragma unroll
                    for (int d = 0; d < N; d++) pn[d] = fma(cdt, k1[d], p0[d]);
                } else if (SCHEME == 1) {
                    const double cdt = (1.0 * alpha) * dt;
#pragma unroll
                    for (int d = 0; d < N; d++) qq[d] = fma(cdt, k1[d], p0[d]);
                    adv_interp<N, UNIFORM, AFFINE, false, INTERP>(g, sm, Vp, c0, r0, gd0, cell1, qq, k2, amask);
                    if (alpha == 0.5) {
#pragma unroll
                        for (int d = 0; d < N; d++) pn[d] = fma(1.0 * dt, k2[d], p0[d]);
                    } else {
                        const double bb = 0.5 * (1.0 / alpha), aa = 1.0 - bb;
#pragma unroll
                        for (int d = 0; d < N; d++) pn[d] = fma(1.0 * dt, fma(bb, k2[d], aa * k1[d]), p0[d]);
                    }
                } else {
                    double k3[3], k4[3];
#pragma unroll
                    for (int d = 0; d < N; d++) qq[d] = p0[d] + dt * k1[d] / 2;
                    adv_interp<N, UNIFORM, AFFINE, false, INTERP>(g, sm, Vp, c0, r0, gd0, cell1, qq, k2, amask);
#pragma unroll
                    for (int d = 0; d < N; d++) qq[d] = p0[d] + dt * k2[d] / 2;
                    adv_interp<N, UNIFORM, AFFINE, false, INTERP>(g, sm, Vp, c0, r0, gd0, cell1, qq, k3, amask);
#pragma unroll
                    for (int d = 0; d < N; d++) qq[d] = p0[d] + dt * k3[d];
                    adv_interp<N, UNIFORM, AFFINE, false, INTERP>(g, sm, Vp, c0, r0, gd0, cell1, qq, k4, amask);
#pragma unroll
                    for (int d = 0; d < N; d++) pn[d] = p0[d] + dt * (((k1[d] + 2 * k2[d]) + 2 * k3[d]) + k4[d]) / 6;
                }
#pragma unroll
                for (int d = 0; d < N; d++) co.p[d][e] = pn[d];
                if (HINT) {
                    const double *xvs = sm + L::XV_OFF;
                    double va[3];
#pragma unroll
                    for (int d = 0; d < N; d++) va[d] = AFFINE ? fma(gd0[d], g.aff_dv[d], g.aff_v0[d]) : xvs[d * L::VEC + r0[d]];
                    const int ci3[3] = {b0[0] + l, cy, cz};
                    int code = g.cls_fast ? jp_classify_fast<N>(g, ci3, va, pn) : -1;
                    if (__any_sync(amask, code < 0) && code < 0) {
                        // within 1e-4 dx of a vertex, far away, NaN / Inf, or a vector grid: the exact comparisons.
                        // The staged vertex segments hold NaN outside the grid, the convention of k_move_classify3.
                        double vm[3], vb[3], vp[3];
#pragma unroll
                        for (int d = 0; d < N; d++) {
                            const double *xv = xvs + d * L::VEC + r0[d];
                            vm[d] = xv[-1]; va[d] = xv[0]; vb[d] = xv[1]; vp[d] = xv[2];
                        }
                        code = jp_classify_particle<N>(g, vm, va, vb, vp, pn);
                    }
                    code_all[(warp * 32 + l) * ADV_HINT_ROW + (cur_l >> 5)] = (uint8_t)code;
                }
            }
        }
        if (!any_next) break;
        cur_valid = nxt_valid; cur_l = nxt_l; cur_e = nxt_e;
#pragma unroll
        for (int d = 0; d < N; d++) cur_p[d] = nxt_p[d];
        __syncwarp();
    }
    if (HINT) {
        // ---- 4. lane = cell: the words k_move_classify3 would have produced (code bytes by slot, eight slots per word)
        __syncwarp();
        const int64_t c = crow + cx;
        const uint64_t *row = reinterpret_cast<const uint64_t *>(code_all + (size_t)(warp * 32 + lane) * ADV_HINT_ROW);
        const uint64_t STAYS = 0x0101010101010101ull * JP_CLS_STAY;
        uint64_t lv = 0;
        unsigned cplx = 0;
        uint64_t codew = 0;
        int k = 0;
        for (int q = 0; q * 8 < g.S; q++) {
            const uint64_t w = row[q];
            const uint64_t x = w ^ STAYS;                                                          // non-zero bytes = leavers
            const uint64_t t = (((x & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | x) & 0x8080808080808080ull;
            unsigned bits = (unsigned)(((t >> 7) * 0x0102040810204080ull) >> 56);                  // bit i = byte i is non-zero
            lv |= (uint64_t)bits << (8 * q);
            while (bits) {                                                                         // this word's leavers, in slot order
                const int i = __ffs((int)bits) - 1;
                bits &= bits - 1;
                int code = (int)((w >> (8 * i)) & 255);
                if (code > JP_CLS_CPLX) { cplx |= 1u << ((code - JP_CLS_CPLX - 1) & 3); code = JP_CODE_DELETE; }
                codew |= (uint64_t)code << (8 * (k & 7));
                if ((++k & 7) == 0) { if (ok) ws.code[(int64_t)((k >> 3) - 1) * g.C + c] = codew; codew = 0; }
            }
        }
        if (ok && (k & 7)) ws.code[(int64_t)(k >> 3) * g.C + c] = codew;
        if (ok) { ws.occ[c] = m; ws.occ0[c] = m; ws.leave[c] = lv; }
        const unsigned wc = __reduce_or_sync(0xffffffffu, cplx);
        if (wc && lane == 0) atomicOr(complex_flag, wc);
    }
}

// true when xvel[c][d] is the vertex vector for c == d and the ghosted-centre vector otherwise
static inline bool jp_standard_staggering(const JpGrid &g) {
    if (!g.fast) return false;
    for (int c = 0; c < g.ndim; c++)
        for (int d = 0; d < g.ndim; d++) {
            if (g.vkind[c][d] != (c == d ? 1 : 2)) return false;
            if (c != d && !g.xg[d]) return false;
        }
    return true;
}
