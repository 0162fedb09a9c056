// jp_emul.cpp -- CPU emulation of the CUDA kernels' decomposition (TEST ONLY).
// Compiles justpic/jl_b200/csrc/jp_core.h (the exact source the kernels use)
// for the host and mimics each kernel's thread decomposition with plain loops:
// classify pass + ordered colour sweeps for move/inject, fast/literal velocity
// interpolation for advect.  Lets the no-GPU test-suite check the kernel logic
// against the oracle.  Never loaded by the product.
// Build: g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -mfma
#include <stdlib.h>
#include <vector>
#include "../../justpic/jl_b200/csrc/jp_host_grid.h"

struct Emul {
    JpGrid g;
    std::vector<double> h;
    std::vector<uint64_t> occ, leave;
    std::vector<uint8_t> flag;
};

extern "C" Emul *jpe_create(const jp_grid_desc *d) {
    Emul *e = new Emul;
    JpGridOffsets off;
    if (jp_grid_build(d, e->g, e->h, off)) { delete e; return nullptr; }
    jp_grid_rebase(e->g, off, e->h.data());
    e->occ.assign(e->g.C, 0); e->leave.assign(e->g.C, 0); e->flag.assign(e->g.C, 0);
    return e;
}
extern "C" void jpe_destroy(Emul *e) { delete e; }
extern "C" int jpe_fast(Emul *e) { return e->g.fast; }
extern "C" void jpe_vkind(Emul *e, int *out) { for (int c = 0; c < 3; c++) for (int d = 0; d < 3; d++) out[c * 3 + d] = e->g.vkind[c][d]; }

template <int N> static void init_t(Emul *e, double *const *co, uint8_t *index, int npq, uint64_t seed) {
    for (int64_t c = 0; c < e->g.C; c++) jp_init_cell<N>(e->g, co, index, npq, seed, c);
}
extern "C" int jpe_init(Emul *e, double *const *co, uint8_t *index, int nxcell, uint64_t seed) {
    const int NQ = e->g.ndim == 2 ? 4 : 8, npq = (nxcell + NQ - 1) / NQ;
    if (npq * NQ > e->g.S) return -1;
    if (e->g.ndim == 2) init_t<2>(e, co, index, npq, seed); else init_t<3>(e, co, index, npq, seed);
    return 0;
}

template <int N, int SCHEME, bool FAST, bool UNIFORM>
static void advect_t(Emul *e, double *const *co, const uint8_t *index, const double *const *V, double alpha, double dt) {
    const JpGrid &g = e->g;
    for (int64_t c = 0; c < g.C; c++) {
        int ci[3];
        jp_cell_ijk<N>(g, c, ci);
        const int cell1[3] = {ci[0] + 1, ci[1] + 1, ci[2] + 1};
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            if (!index[el]) continue;
            double p0[3], p1[3];
            for (int d = 0; d < N; d++) p0[d] = co[d][el];
            jp_advect_particle<N, SCHEME, FAST, UNIFORM>(g, alpha, V, dt, cell1, p0, p1);
            for (int d = 0; d < N; d++) co[d][el] = p1[d];
        }
    }
}
template <int N, int SCHEME>
static void advect_s(Emul *e, double *const *co, const uint8_t *index, const double *const *V, double alpha, double dt, int force_literal) {
    if (e->g.fast && !force_literal) {
        if (e->g.uniform) advect_t<N, SCHEME, true, true>(e, co, index, V, alpha, dt);
        else advect_t<N, SCHEME, true, false>(e, co, index, V, alpha, dt);
    } else advect_t<N, SCHEME, false, false>(e, co, index, V, alpha, dt);
}
extern "C" int jpe_advect(Emul *e, double *const *co, const uint8_t *index, int scheme, double alpha, const double *const *V, double dt, int force_literal) {
    if (e->g.ndim == 2) {
        if (scheme == 0) advect_s<2, 0>(e, co, index, V, alpha, dt, force_literal);
        else if (scheme == 1) advect_s<2, 1>(e, co, index, V, alpha, dt, force_literal);
        else advect_s<2, 2>(e, co, index, V, alpha, dt, force_literal);
    } else {
        if (scheme == 0) advect_s<3, 0>(e, co, index, V, alpha, dt, force_literal);
        else if (scheme == 1) advect_s<3, 1>(e, co, index, V, alpha, dt, force_literal);
        else advect_s<3, 2>(e, co, index, V, alpha, dt, force_literal);
    }
    return 0;
}

template <int N>
static void move_t(Emul *e, double *const *co, uint8_t *index, const JpArgs &args, int64_t *stats) {
    const JpGrid &g = e->g;
    if (g.S > JP_MAX_SLOTS) {                // wide cells: k_move_sweep_wide, literal slot loop on the index bytes
        int st[3] = {0, 0, 0};
        for (int ox = 0; ox < 3; ox++)
            for (int oy = 0; oy < 3; oy++)
                for (int oz = 0; oz < (N == 3 ? 3 : 1); oz++)
                    for (int k = oz; k < g.n[2]; k += 3)
                        for (int j = oy; j < g.n[1]; j += 3)
                            for (int i = ox; i < g.n[0]; i += 3) {
                                int ci[3] = {i, j, k};
                                jp_move_cell_wide<N>(g, co, index, args, jp_cell_lin<N>(g, ci), ci, st, false);
                            }
        stats[0] = st[0]; stats[1] = st[1]; stats[2] = st[2];
        return;
    }
    // pass A: classify (k_move_classify)
    for (int64_t c = 0; c < g.C; c++) {
        int ci[3];
        jp_cell_ijk<N>(g, c, ci);
        uint64_t m = 0, lv = 0;
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            if (!index[el]) continue;
            m |= 1ull << s;
            double p[3];
            for (int d = 0; d < N; d++) p[d] = co[d][el];
            if (jp_move_leaves<N>(g, ci, p)) lv |= 1ull << s;
        }
        e->occ[c] = m; e->leave[c] = lv;
    }
    // pass B: 3^N ordered colour sweeps (k_move_sweep)
    int st[3] = {0, 0, 0};
    for (int ox = 0; ox < 3; ox++)
        for (int oy = 0; oy < 3; oy++)
            for (int oz = 0; oz < (N == 3 ? 3 : 1); oz++)
                for (int k = oz; k < g.n[2]; k += 3)
                    for (int j = oy; j < g.n[1]; j += 3)
                        for (int i = ox; i < g.n[0]; i += 3) {
                            int ci[3] = {i, j, k};
                            jp_move_cell<N>(g, co, index, args, e->occ.data(), e->leave.data(), jp_cell_lin<N>(g, ci), ci, st);
                        }
    stats[0] = st[0]; stats[1] = st[1]; stats[2] = st[2];
}
extern "C" int jpe_move(Emul *e, double *const *co, uint8_t *index, double *const *args, int nargs, int64_t *stats) {
    JpArgs a; a.n = nargs;
    for (int i = 0; i < nargs; i++) a.a[i] = args[i];
    if (e->g.ndim == 2) move_t<2>(e, co, index, a, stats); else move_t<3>(e, co, index, a, stats);
    return 0;
}

template <int N>
static int64_t inject_t(Emul *e, double *const *co, uint8_t *index, const JpArgs &args, int min_xcell, uint64_t seed, uint32_t step) {
    const JpGrid &g = e->g;
    const int NQ = N == 2 ? 4 : 8, min_xq = (min_xcell + NQ - 1) / NQ;
    for (int64_t c = 0; c < g.C; c++) {      // k_inject_classify (wide cells, S > 64: k_inject_wide visits every cell)
        if (g.S > JP_MAX_SLOTS) { e->flag[c] = 1; continue; }
        int ci[3];
        jp_cell_ijk<N>(g, c, ci);
        double vq[3], dq[3];
        for (int d = 0; d < N; d++) { vq[d] = g.xv[d][ci[d]]; dq[d] = jp_d_of(g.xv[d], g.uniform, ci[d]) / 2; }
        int nq0 = 0, nlive = 0;
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            if (!index[el]) continue;
            nlive++;
            double p[3];
            for (int d = 0; d < N; d++) p[d] = co[d][el];
            nq0 += jp_isincell<N>(p, vq, dq) ? 1 : 0;
        }
        e->flag[c] = jp_inject_candidate(nq0, nlive, g.S, min_xq) ? 1 : 0;
    }
    int64_t inj = 0;
    for (int ox = 0; ox < 2; ox++)           // k_inject_sweep
        for (int oy = 0; oy < 2; oy++)
            for (int oz = 0; oz < (N == 3 ? 2 : 1); oz++)
                for (int k = oz; k < g.n[2]; k += 2)
                    for (int j = oy; j < g.n[1]; j += 2)
                        for (int i = ox; i < g.n[0]; i += 2) {
                            int ci[3] = {i, j, k};
                            const int64_t c = jp_cell_lin<N>(g, ci);
                            if (!e->flag[c]) continue;
                            inj += jp_inject_cell<N>(g, co, index, args, min_xcell, seed, step, c, ci);
                        }
    return inj;
}
extern "C" int64_t jpe_inject(Emul *e, double *const *co, uint8_t *index, double *const *args, int nargs, int min_xcell, uint64_t seed, uint32_t step) {
    JpArgs a; a.n = nargs;
    for (int i = 0; i < nargs; i++) a.a[i] = args[i];
    return e->g.ndim == 2 ? inject_t<2>(e, co, index, a, min_xcell, seed, step) : inject_t<3>(e, co, index, a, min_xcell, seed, step);
}


// ---- move_particles! classification of one particle, three ways (tests/test_classify_emul.py):
// out[0] = jp_classify_fast (single-precision pre-filter; -2 when the grid does not qualify),
// out[1] = jp_classify_particle with the four vertices around the storage cell as k_move_classify3 / the
//          advection hand-off load them (NaN outside the grid),
// out[2] = the literal route of move_kernel! (src/Particles/move_safe.jl:86-106): isincell -> indomain -> bisection,
//          expressed in the same codes (JP_CLS_CPLX + r where the planner would hand over to the direct sweeps).
template <int N> static void classify_t(Emul *e, const int *ci, const double *p, int *out) {
    const JpGrid &g = e->g;
    double am[3], a[3], b[3], bp[3], corner[3], dx[3];
    for (int d = 0; d < N; d++) {
        am[d] = ci[d] > 0 ? g.xv[d][ci[d] - 1] : NAN;
        a[d] = g.xv[d][ci[d]];
        b[d] = g.xv[d][ci[d] + 1];
        bp[d] = ci[d] + 2 <= g.n[d] ? g.xv[d][ci[d] + 2] : NAN;
        corner[d] = a[d]; dx[d] = jp_d_of(g.xv[d], g.uniform, ci[d]);
    }
    out[0] = g.cls_fast ? jp_classify_fast<N>(g, ci, a, p) : -2;
    out[1] = jp_classify_particle<N>(g, am, a, b, bp, p);
    int lit;
    if (jp_isincell<N>(p, corner, dx)) lit = JP_CLS_STAY;
    else {
        bool indom = true;
        for (int d = 0; d < N; d++) indom = indom && (g.xv[d][0] < p[d] && p[d] < g.xv[d][g.n[d]]);
        if (!indom) lit = JP_CODE_DELETE;
        else {
            int dv[3] = {0, 0, 0}, nc[3] = {0, 0, 0};
            bool far = false;
            double c2[3], dx2[3];
            for (int d = 0; d < N; d++) {
                nc[d] = jp_bisect1(p[d], g.xv[d], g.n[d] + 1, ci[d] + 1) - 1;
                dv[d] = nc[d] - ci[d];
                far = far || dv[d] < -1 || dv[d] > 1;
                c2[d] = g.xv[d][nc[d]]; dx2[d] = jp_d_of(g.xv[d], g.uniform, nc[d]);
            }
            if (far) lit = JP_CLS_CPLX + 1;
            else if (dv[0] == 0 && dv[1] == 0 && dv[2] == 0) lit = JP_CLS_CPLX + 2;
            else if (!jp_isincell<N>(p, c2, dx2)) lit = JP_CLS_CPLX + 3;
            else lit = (dv[0] + 1) + 3 * (dv[1] + 1) + (N == 3 ? 9 * (dv[2] + 1) : 9);
        }
    }
    out[2] = lit;
}
extern "C" int jpe_classify(Emul *e, const int *ci, const double *p, int *out) {
    if (e->g.ndim == 2) classify_t<2>(e, ci, p, out); else classify_t<3>(e, ci, p, out);
    return e->g.cls_fast;
}

// ---- cell-local interpolation / cleanup kernels: the kernels' loops (k_g2p, k_c2p, k_p2g, k_p2c, k_phase, k_clean,
// k_advect_hi in justpic_sm100a.cu) around the SAME per-particle functions of jp_core.h
template <int N> static void g2p_t(Emul *e, double *const *co, const uint8_t *index, double *Fp, const double *F) {
    const JpGrid &g = e->g;
    for (int64_t c = 0; c < g.C; c++) {
        int ci[3]; jp_cell_ijk<N>(g, c, ci);
        double v[8], xcorner[3], idx[3];
        const int64_t s1 = g.n[0] + 1, s2 = (int64_t)(g.n[0] + 1) * (g.n[1] + 1);
        jp_corners<N>(F, ci[0] + s1 * ci[1] + (N == 3 ? s2 * ci[2] : 0), s1, s2, v);
        for (int d = 0; d < N; d++) { xcorner[d] = g.xv[d][ci[d]]; idx[d] = 1.0 / jp_d_of(g.xv[d], g.uniform, ci[d]); }
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            if (!index[el]) continue;
            double p[3];
            for (int d = 0; d < N; d++) p[d] = co[d][el];
            Fp[el] = jp_g2p<N>(v, xcorner, idx, p);
        }
    }
}
extern "C" void jpe_grid2particle(Emul *e, double *const *co, const uint8_t *index, double *Fp, const double *F) {
    if (e->g.ndim == 2) g2p_t<2>(e, co, index, Fp, F); else g2p_t<3>(e, co, index, Fp, F);
}

template <int N> static void c2p_t(Emul *e, double *const *co, double *Fp, const double *Fc) {
    const JpGrid &g = e->g;
    for (int64_t c = 0; c < g.C; c++) {
        int ci[3]; jp_cell_ijk<N>(g, c, ci);
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            double p[3]; bool nan = false;
            for (int d = 0; d < N; d++) { p[d] = co[d][el]; nan |= std::isnan(p[d]); }
            if (nan) continue;
            Fp[el] = jp_c2p<N>(g, Fc, ci, p);
        }
    }
}
extern "C" void jpe_centroid2particle(Emul *e, double *const *co, double *Fp, const double *Fc) {
    if (e->g.ndim == 2) c2p_t<2>(e, co, Fp, Fc); else c2p_t<3>(e, co, Fp, Fc);
}

template <int N> static void p2g_t(Emul *e, double *const *co, const uint8_t *index, double *F, const double *Fp) {
    const JpGrid &g = e->g;
    const int nx = g.n[0], ny = g.n[1], nz = N == 3 ? g.n[2] : 1;
    for (int kn = 0; kn <= (N == 3 ? nz : 0); kn++)
        for (int jn = 0; jn <= ny; jn++)
            for (int in = 0; in <= nx; in++) {
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
                                const int64_t el = c + (int64_t)s * g.C;
                                if (!index[el]) continue;
                                double p[3];
                                for (int d = 0; d < N; d++) p[d] = co[d][el];
                                const double wi = jp_p2g_weight<N>(xn, p);
                                w += wi;
                                wF = fma(wi, Fp[el], wF);
                            }
                        }
                    }
                }
                const int64_t nd = in + (int64_t)(nx + 1) * (jn + (N == 3 ? (int64_t)(ny + 1) * kn : 0));
                F[nd] = N == 2 ? wF / w : wF * (1.0 / w);
            }
}
extern "C" void jpe_particle2grid(Emul *e, double *const *co, const uint8_t *index, double *F, const double *Fp) {
    if (e->g.ndim == 2) p2g_t<2>(e, co, index, F, Fp); else p2g_t<3>(e, co, index, F, Fp);
}

template <int N> static void p2c_t(Emul *e, double *const *co, double *Fc, const double *Fp) {
    const JpGrid &g = e->g;
    for (int64_t c = 0; c < g.C; c++) {
        int ci[3]; jp_cell_ijk<N>(g, c, ci);
        double xcn[3], idi[3];
        for (int d = 0; d < N; d++) { xcn[d] = g.xc[d][ci[d]]; idi[d] = 1.0 / jp_d_of(g.xv[d], g.uniform, ci[d]); }
        double w = 0.0, wF = 0.0;
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            double p[3];
            for (int d = 0; d < N; d++) p[d] = co[d][el];
            if (N == 2 ? (std::isnan(p[0]) || std::isnan(p[1])) : std::isnan(p[0])) continue;
            const double wi = jp_bilinear_weight<N>(xcn, p, idi);
            w += wi;
            wF = fma(wi, Fp[el], wF);
        }
        Fc[c] = N == 2 ? wF / w : wF * (1.0 / w);
    }
}
extern "C" void jpe_particle2centroid(Emul *e, double *const *co, double *Fc, const double *Fp) {
    if (e->g.ndim == 2) p2c_t<2>(e, co, Fc, Fp); else p2c_t<3>(e, co, Fc, Fp);
}

template <int N> static void phase_t(Emul *e, double *const *co, double *ratios, const double *phases, int K) {
    const JpGrid &g = e->g;
    for (int64_t c = 0; c < g.C; c++) {
        int ci[3]; jp_cell_ijk<N>(g, c, ci);
        double xcn[3], idi[3], w[JP_MAX_PHASES];
        for (int d = 0; d < N; d++) { xcn[d] = g.xc[d][ci[d]]; idi[d] = 1.0 / jp_d_of(g.xv[d], g.uniform, ci[d]); }
        for (int k = 0; k < K; k++) w[k] = 0.0;
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            if (std::isnan(co[0][el])) continue;
            const double p[3] = {co[0][el], co[1][el], N == 3 ? co[2][el] : 0.0};
            const double x = jp_bilinear_weight<N>(xcn, p, idi);
            for (int k = 0; k < K; k++) w[k] = w[k] + (phases[el] == (double)(k + 1) ? x : copysign(0.0, x));
        }
        double sum = w[0];
        for (int k = 1; k < K; k++) sum = sum + w[k];
        const double inv = 1.0 / sum;
        for (int k = 0; k < K; k++) ratios[c + (int64_t)k * g.C] = w[k] * inv;
    }
}
extern "C" void jpe_phase_ratios_center(Emul *e, double *const *co, double *ratios, const double *phases, int K) {
    if (e->g.ndim == 2) phase_t<2>(e, co, ratios, phases, K); else phase_t<3>(e, co, ratios, phases, K);
}

template <int N> static void clean_t(Emul *e, double *const *co, uint8_t *index, const JpArgs &args) {
    const JpGrid &g = e->g;
    for (int64_t c = 0; c < g.C; c++) {
        int ci[3]; jp_cell_ijk<N>(g, c, ci);
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            if (!index[el]) continue;
            double p[3];
            for (int d = 0; d < N; d++) p[d] = co[d][el];
            if (jp_clean_removes<N>(g, ci, p)) {
                index[el] = 0;
                for (int d = 0; d < N; d++) co[d][el] = NAN;
                for (int a = 0; a < args.n; a++) args.a[a][el] = NAN;
            }
        }
    }
}
extern "C" void jpe_clean(Emul *e, double *const *co, uint8_t *index, double *const *args, int nargs) {
    JpArgs a; a.n = nargs;
    for (int i = 0; i < nargs; i++) a.a[i] = args[i];
    if (e->g.ndim == 2) clean_t<2>(e, co, index, a); else clean_t<3>(e, co, index, a);
}

template <int N, int SCHEME, int INTERP>
static void advect_hi_t(Emul *e, double *const *co, const uint8_t *index, const double *const *V, double alpha, double dt) {
    const JpGrid &g = e->g;
    for (int64_t c = 0; c < g.C; c++) {
        int ci[3]; jp_cell_ijk<N>(g, c, ci);
        const int cell1[3] = {ci[0] + 1, ci[1] + 1, ci[2] + 1};
        for (int s = 0; s < g.S; s++) {
            const int64_t el = c + (int64_t)s * g.C;
            if (!index[el]) continue;
            double p0[3], p1[3];
            for (int d = 0; d < N; d++) p0[d] = co[d][el];
            jp_advect_particle_hi<N, SCHEME, INTERP>(g, alpha, V, dt, cell1, p0, p1);
            for (int d = 0; d < N; d++) co[d][el] = p1[d];
        }
    }
}
template <int N, int INTERP>
static void advect_hi_s(Emul *e, double *const *co, const uint8_t *index, int scheme, const double *const *V, double alpha, double dt) {
    if (scheme == 0) advect_hi_t<N, 0, INTERP>(e, co, index, V, alpha, dt);
    else if (scheme == 1) advect_hi_t<N, 1, INTERP>(e, co, index, V, alpha, dt);
    else advect_hi_t<N, 2, INTERP>(e, co, index, V, alpha, dt);
}
extern "C" void jpe_advect_interp(Emul *e, double *const *co, const uint8_t *index, int scheme, double alpha, const double *const *V, double dt, int interp) {
    if (e->g.ndim == 2) { if (interp == 1) advect_hi_s<2, 1>(e, co, index, scheme, V, alpha, dt); else advect_hi_s<2, 2>(e, co, index, scheme, V, alpha, dt); }
    else                { if (interp == 1) advect_hi_s<3, 1>(e, co, index, scheme, V, alpha, dt); else advect_hi_s<3, 2>(e, co, index, scheme, V, alpha, dt); }
}
