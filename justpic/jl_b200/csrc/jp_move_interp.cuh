// jp_move_interp.cuh -- move_particles! -> particle2grid! / phase_ratios_center! hand-off (JP_OPT_MOVE_INTERP).
//
// In the reference's time loop (scripts/temperature_advection3D.jl:69-75, test/test_2D.jl:517-524) move_particles! is
// followed by particle2grid!(T, pT, particles) and, in the real consumer, phase_ratios_center!(phase_ratios, particles,
// phases): both stream every particle again (coordinates + one field each: 34.3 + 34.7 B / particle)
// although the scatter pass of the move has just had ~77 % of those sectors in its hands (at a migrant fraction of
// 0.38 nearly every 32-byte sector of every array holds a changed slot and is fetched as fill for the partial write).
//
// k_move_scatter_interp is k_move_scatter (jp_move_plan.cuh, pass E) that also reads the live slots it does not
// change -- coordinates, the particle2grid! field and the phase field only -- and accumulates, in slot order,
//   * the per-cell partial sums of the two-pass particle2grid! (exactly what k_p2g_cell computes:
//     src/Interpolations/particle_to_grid.jl:113-151 restated cell-centrically, see justpic_sm100a.cu), and
//   * the centre phase ratios (exactly k_phase: src/PhaseRatios/centers.jl:13-30, utils.jl:45-74)
// of the cell's FINAL content.  jp_particle2grid then only runs its node pass and jp_phase_ratios_center copies the
// ratios out.  Same arithmetic in the same order as the stand-alone kernels, so results are bit-identical to them
// (tests/test_gpu_interp_handoff.py); liveness is the final occupancy word instead of the isnan(px) probe of
// phase_ratios_center! -- identical for every container this library (or the reference) produces, where dead slots
// hold NaN.
#pragma once
#include <type_traits>

// one particle into the 2^N corner sums of its cell (the body of k_p2g_cell)
template <int N, bool FASTW>
__device__ __forceinline__ void jp_p2g_cell_accum(const double (&xn)[3][2], const double *p, double f, double *aw, double *awf) {
    constexpr int NQ = N == 2 ? 4 : 8;
    double d2[3][2];
#pragma unroll
    for (int d = 0; d < N; d++) {
        const double a0 = xn[d][0] - p[d], a1 = xn[d][1] - p[d];
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
        awf[q] = fma(wi, f, awf[q]);
    }
}

// one particle into the centre phase weights of its cell (the body of k_phase)
template <int N, int KMAX>
__device__ __forceinline__ void jp_phase_accum(const double *xcn, const double *idi, const double *p, double ph, int K, double *w) {
    const double x = jp_bilinear_weight<N>(xcn, p, idi);
#pragma unroll
    for (int k = 0; k < KMAX; k++)
        if (k < K) w[k] = w[k] + (ph == (double)(k + 1) ? x : copysign(0.0, x));
}

#ifndef JP_MINB_SCATTER_INTERP
#define JP_MINB_SCATTER_INTERP 2
#endif
#ifndef JP_SCI_U
#define JP_SCI_U 2          // slots per load batch
#endif

struct MoveInterp {
    int iT, iP;             // index into MoveArrays of the particle2grid! field / the phase field (-1: not requested)
    int K;                  // number of phases
    double *PW, *PWF;       // [2^N][C] two-pass particle2grid! partial sums
    double *RC;             // [K][C] centre phase ratios
};

template <int N, int KMAX, bool FASTW>
__global__ void __launch_bounds__(256, JP_MINB_SCATTER_INTERP) k_move_scatter_interp(JpGrid g, MovePlanWs ws, MoveArrays arrs, uint8_t *index,
                                                                                     const double *__restrict__ stage, MoveInterp mi,
                                                                                     const unsigned int *__restrict__ skip_flag) {
    constexpr int NQ = N == 2 ? 4 : 8;
    constexpr int U = JP_SCI_U;
    if (*skip_flag) return;                                   // the call takes the direct sweeps (decided on the device)
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const uint64_t amask = ok ? ws.arrmask[c] : 0, lmask = ok ? ws.leave[c] : 0, occf = ok ? ws.occ[c] : 0;
    const uint64_t changed = amask | lmask;
    const uint64_t visit = changed | occf;
    const int64_t base = ok ? ws.off[c] : 0;
    const bool do_p2g = mi.iT >= 0, do_ph = mi.iP >= 0;
    double xn[3][2], xcn[3], idi[3];
    if (ok) {
#pragma unroll
        for (int d = 0; d < N; d++) {
            xn[d][0] = g.xv[d][ci[d]]; xn[d][1] = g.xv[d][ci[d] + 1];
            xcn[d] = g.xc[d][ci[d]]; idi[d] = 1.0 / jp_d_of(g.xv[d], g.uniform, ci[d]);
        }
    }
    double aw[NQ], awf[NQ], w[KMAX];
#pragma unroll
    for (int q = 0; q < NQ; q++) { aw[q] = 0.0; awf[q] = 0.0; }
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = 0.0;
    const int AS = (arrs.n + 3) & ~3;
    for (int s0 = 0; s0 < g.S; s0 += U) {
        const unsigned vb = (unsigned)(visit >> s0) & ((1u << U) - 1u);
        if (!__any_sync(0xffffffffu, vb != 0)) continue;
        const unsigned chb = (unsigned)(changed >> s0) & ((1u << U) - 1u);
        const unsigned arb = (unsigned)(amask >> s0) & ((1u << U) - 1u);
        const unsigned fb = (unsigned)(occf >> s0) & ((1u << U) - 1u);
        int64_t pos[U];
#pragma unroll
        for (int u = 0; u < U; u++) pos[u] = (base + __popcll(amask & ((1ull << (s0 + u)) - 1))) * AS;
        double cap[U][5];                                      // x, y, z, particle2grid! field, phase of the slot's final occupant
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int q = 0; q < 5; q++) cap[u][q] = 0.0;
        // arrays in register batches of JP_MV_A; the first batch (coordinates + the first field) has compile-time indices
        auto batch = [&](const int a0, auto first) {
            constexpr bool FIRST = decltype(first)::value;
            double v[U][JP_MV_A];
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int a = 0; a < JP_MV_A; a++) {
                    const int ia = a0 + a;
                    v[u][a] = NAN;
                    if (ia < arrs.n) {
                        if ((chb >> u) & 1u) { if ((arb >> u) & 1u) v[u][a] = stage[pos[u] + ia]; }
                        else if (((fb >> u) & 1u) && ((FIRST && a < N) || ia == mi.iT || ia == mi.iP))
                            v[u][a] = arrs.a[ia][c + (int64_t)(s0 + u) * g.C];
                    }
                }
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int a = 0; a < JP_MV_A; a++) {
                    const int ia = a0 + a;
                    if (ia < arrs.n && ((chb >> u) & 1u)) arrs.a[ia][c + (int64_t)(s0 + u) * g.C] = v[u][a];
                    if (FIRST && a < N) cap[u][a] = v[u][a];
                    if (ia == mi.iT) cap[u][3] = v[u][a];
                    if (ia == mi.iP) cap[u][4] = v[u][a];
                }
        };
        batch(0, std::true_type());
        for (int a0 = JP_MV_A; a0 < arrs.n; a0 += JP_MV_A) batch(a0, std::false_type());
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int s = s0 + u;
            if ((chb >> u) & 1u) {
                if ((arb >> u) & 1u) { if (!((lmask >> s) & 1ull)) index[c + (int64_t)s * g.C] = 1; }
                else index[c + (int64_t)s * g.C] = 0;
            }
            if ((fb >> u) & 1u) {                              // slot order = the reference's summation order
                if (do_p2g) jp_p2g_cell_accum<N, FASTW>(xn, cap[u], cap[u][3], aw, awf);
                if (do_ph) jp_phase_accum<N, KMAX>(xcn, idi, cap[u], cap[u][4], mi.K, w);
            }
        }
    }
    if (ok) {
        if (do_p2g) {
#pragma unroll
            for (int q = 0; q < NQ; q++) { mi.PW[(int64_t)q * g.C + c] = aw[q]; mi.PWF[(int64_t)q * g.C + c] = awf[q]; }
        }
        if (do_ph) {
            double sum = w[0];
#pragma unroll
            for (int k = 1; k < KMAX; k++) if (k < mi.K) sum = sum + w[k];
            const double inv = 1.0 / sum;
#pragma unroll
            for (int k = 0; k < KMAX; k++) if (k < mi.K) mi.RC[c + (int64_t)k * g.C] = w[k] * inv;
        }
    }
}


// ---- the same pass for the usual argument order -- move_particles!(particles, (Fp, phases, ...)): the particle2grid! field is
// the first particle field and the phase field (if registered) the second -- with 4-slot load batches like k_move_scatter:
// coordinates + Fp (3-D) / coordinates + Fp + phases (2-D) are exactly the first register batch of JP_MV_A = 4 arrays, so what
// the scatter has in registers is what the accumulation needs, with compile-time indices throughout.
// Measured at 256^3 (B200, profiles/r02c_ab_scatter.log; the separate kernels take 11.9 + 5.7 + 3.9 = 21.5 ms):
//   generic kernel above (run-time indices, 2-slot batches)            20.0 ms
//   this kernel, 2-slot batches, sums in registers (128 regs, 2 CTA/SM) 14.9 ms   <- shipped
//   2-slot batches, sums in shared memory                              15.2 ms
//   4-slot batches (spills at 128 registers), sums in smem / registers 16.8 / 16.9 ms
//   cell constants in shared memory as well                            19.6 ms;  3 CTAs/SM at 80 registers (spills): 37 ms
#ifndef JP_SCI_FAST_U
#define JP_SCI_FAST_U 2
#endif
#ifndef JP_SCI_PREFETCH
#define JP_SCI_PREFETCH 2        // batches ahead whose sectors are requested with prefetch.global.L1 (r02n: 14.9 -> 14.0 ms; 1: 14.0, 4: 14.4)
#endif
template <int N, int KMAX, bool FASTW, bool HAS_PH>
__global__ void __launch_bounds__(256, JP_MINB_SCATTER_INTERP) k_move_scatter_interp_fast(JpGrid g, MovePlanWs ws, MoveArrays arrs, uint8_t *index,
                                                                                          const double *__restrict__ stage, MoveInterp mi,
                                                                                          const unsigned int *__restrict__ skip_flag) {
    constexpr int NQ = N == 2 ? 4 : 8;
    constexpr int U = JP_SCI_FAST_U;
    constexpr int IT = N;                                     // arrays: x, y[, z], Fp, phases, others...
    constexpr int IP = N + 1;
    constexpr int NG = HAS_PH ? N + 2 : N + 1;                // arrays the accumulation reads
    if (*skip_flag) return;
    int ci[3]; int64_t c;
    const bool ok = tile_cell<N>(g, ci, c);
    const int tid = threadIdx.y * JP_BX + threadIdx.x;
    (void)tid;
    const uint64_t amask = ok ? ws.arrmask[c] : 0, lmask = ok ? ws.leave[c] : 0, occf = ok ? ws.occ[c] : 0;
    const uint64_t changed = amask | lmask;
    const uint64_t visit = changed | occf;
    const unsigned base = ok ? ws.off[c] : 0;
    double xn[3][2], xcn[3], idi[3];
    if (ok) {
#pragma unroll
        for (int d = 0; d < N; d++) {
            xn[d][0] = g.xv[d][ci[d]]; xn[d][1] = g.xv[d][ci[d] + 1];
            if (HAS_PH) { xcn[d] = g.xc[d][ci[d]]; idi[d] = 1.0 / jp_d_of(g.xv[d], g.uniform, ci[d]); }
        }
    }
    double aw[NQ], awf[NQ];
#pragma unroll
    for (int q = 0; q < NQ; q++) { aw[q] = 0.0; awf[q] = 0.0; }
    double w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; k++) w[k] = 0.0;
    const int AS = (arrs.n + 3) & ~3;
    const double *a0p = arrs.a[0], *a1p = arrs.a[1], *a2p = arrs.a[2], *a3p = arrs.a[3];
    for (int s0 = 0; s0 < g.S; s0 += U) {
        const unsigned vb = (unsigned)(visit >> s0) & ((1u << U) - 1u);
        if (!__any_sync(0xffffffffu, vb != 0)) continue;
        const unsigned chb = (unsigned)(changed >> s0) & ((1u << U) - 1u);
        const unsigned arb = (unsigned)(amask >> s0) & ((1u << U) - 1u);
        const unsigned fb = (unsigned)(occf >> s0) & ((1u << U) - 1u);
        const double *sp[U];                                   // staging record of the slot's arrival
#pragma unroll
        for (int u = 0; u < U; u++) sp[u] = stage + (size_t)(base + (unsigned)__popcll(amask & ((1ull << (s0 + u)) - 1))) * AS;
#if JP_SCI_PREFETCH
        // the sectors of the batch JP_SCI_PREFETCH batches ahead: requested now, no register held, in flight during this batch's arithmetic
        {
            const int sn = s0 + JP_SCI_PREFETCH * U;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int s = sn + u;
                if (s < g.S) {
                    const bool arn = (amask >> s) & 1ull, keepn = !((changed >> s) & 1ull) && ((occf >> s) & 1ull);
                    if (arn) {
                        const double *spn = stage + (size_t)(base + (unsigned)__popcll(amask & ((1ull << s) - 1))) * AS;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(spn));
                        if (arrs.n > 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(spn + 4));
                    } else if (keepn) {
                        const int64_t en = c + (int64_t)s * g.C;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a0p + en));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a1p + en));
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a2p + en));
                        if (NG > 3) asm volatile("prefetch.global.L1 [%0];" ::"l"(a3p + en));
                        if (HAS_PH && N == 3) asm volatile("prefetch.global.L1 [%0];" ::"l"(arrs.a[IP] + en));
                    }
                }
            }
        }
#endif
        double v[U][JP_MV_A], phv[U];
        // ---- loads: arrivals from staging, unchanged occupants from the arrays (only what the accumulation reads)
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t e = c + (int64_t)(s0 + u) * g.C;
            const bool ch = (chb >> u) & 1u, ar = (arb >> u) & 1u, keep = !ch && ((fb >> u) & 1u);
#pragma unroll
            for (int a = 0; a < JP_MV_A; a++) {
                v[u][a] = NAN;
                if (a < arrs.n) {
                    if (ar) v[u][a] = sp[u][a];
                    else if (keep && a < NG) v[u][a] = (a == 0 ? a0p : a == 1 ? a1p : a == 2 ? a2p : a3p)[e];
                }
            }
            phv[u] = NAN;
            if (HAS_PH && N == 3) {
                if (ar) phv[u] = sp[u][IP];
                else if (keep) phv[u] = arrs.a[IP][e];
            }
        }
        // ---- stores of the changed slots
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!((chb >> u) & 1u)) continue;
            const int64_t e = c + (int64_t)(s0 + u) * g.C;
#pragma unroll
            for (int a = 0; a < JP_MV_A; a++) if (a < arrs.n) arrs.a[a][e] = v[u][a];
            if (HAS_PH && N == 3) arrs.a[IP][e] = phv[u];
        }
        // ---- the other particle fields: scatter only
        for (int a0 = (HAS_PH && N == 3) ? JP_MV_A + 1 : JP_MV_A; a0 < arrs.n; a0++) {
            double o[U];
#pragma unroll
            for (int u = 0; u < U; u++) o[u] = (((chb & arb) >> u) & 1u) ? sp[u][a0] : NAN;
#pragma unroll
            for (int u = 0; u < U; u++) if ((chb >> u) & 1u) arrs.a[a0][c + (int64_t)(s0 + u) * g.C] = o[u];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int s = s0 + u;
            if ((chb >> u) & 1u) {
                if ((arb >> u) & 1u) { if (!((lmask >> s) & 1ull)) index[c + (int64_t)s * g.C] = 1; }
                else index[c + (int64_t)s * g.C] = 0;
            }
            if ((fb >> u) & 1u) {                              // slot order = the reference's summation order
                const double p[3] = {v[u][0], v[u][1], N == 3 ? v[u][2] : 0.0};
                const double f = v[u][IT];
                jp_p2g_cell_accum<N, FASTW>(xn, p, f, aw, awf);
                if (HAS_PH) jp_phase_accum<N, KMAX>(xcn, idi, p, N == 3 ? phv[u] : v[u][IP], mi.K, w);
            }
        }
    }
    if (ok) {
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            mi.PW[(int64_t)q * g.C + c] = aw[q]; mi.PWF[(int64_t)q * g.C + c] = awf[q];
        }
        if (HAS_PH) {
            double sum = w[0];
#pragma unroll
            for (int k = 1; k < KMAX; k++) if (k < mi.K) sum = sum + w[k];
            const double inv = 1.0 / sum;
#pragma unroll
            for (int k = 0; k < KMAX; k++) if (k < mi.K) mi.RC[c + (int64_t)k * g.C] = w[k] * inv;
        }
    }
}
