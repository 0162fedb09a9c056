/*
 * justpic_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the JustPIC.jl particle-in-cell hot path, written
 * from the reference's Julia sources (file:line cited at every function; paths
 * are relative to /root/reference).  It exists to check the CUDA library
 * (libjustpic_sm100a.so) bit-for-bit.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * path never does.
 *
 * PARITY PIN STATUS.  Julia is absent from this image, so the reference itself
 * cannot be run.  Pinned by the reference's own known-answer tests (see
 * tests/test_oracle_kat.py): integrator stage formulas
 * (test/test_integrators.jl:11-78) and N-linear lerp
 * (test/test_interpolation_kernels.jl:47-60), plus the reference's property
 * tests (linear-field grid2particle == coordinate, phase ratios sum to 1, cell
 * bracket), and force_injection! by the exact expectations of its reference
 * tests (test/test_2D.jl:301-384, test/test_3D.jl:279-333, transcribed in
 * tests/test_force_injection.py).  move_particles!/inject_particles!/particle2grid! slot-level
 * results are NOT pinned by any reference test ("parity unpinned"): this
 * literal restatement is the only pin.  RNG streams are unpinned in the
 * reference (backend rand()); both this oracle and the CUDA library use
 * Philox4x32-10 keyed (seed, step, cell, slot).
 *
 * Floating-point contract: compile with -O2 -ffp-contract=off; fma() is
 * called exactly where the reference has muladd/fma/@muladd; everything else
 * is unfused, evaluated in the reference's order.
 *
 * Layout: CUDA CellArray layout, element (cell c, slot
 * s) at c + s*C with c = i + nx*(j + ny*k) (0-based), index = 1 byte/slot.
 * Grids: xv[d] (n_d+1 vertices), xc[d] (n_d centres), xvel[comp][dim]
 * staggered velocity grid vectors with ghost nodes.  Velocity component
 * arrays are column-major with extents nvel[comp][0..2].
 *
 * Multi-threading: OpenMP over cells (cell-local kernels) or over same-colour
 * cells (move/inject), exactly the reference's parallel decomposition.  With
 * jpo_set_threads(1) execution is serial and deterministic for any input.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t ndim;            /* 2 or 3 */
    int32_t n[3];            /* cells per dim (n[2] = 1 in 2D) */
    int32_t S;               /* slots per cell (max_xcell) */
    int32_t uniform;         /* 1: scalar spacings x[1]-x[0] (range grids), 0: diff(x)[i] */
    const double *xv[3];     /* n+1 */
    const double *xc[3];     /* n */
    const double *xvel[3][3];/* xvel[comp][dim] */
    int32_t nvel[3][3];      /* lengths of xvel[comp][dim] */
} jpo_grid;

static int g_threads = 1;
void jpo_set_threads(int n) {
    g_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
    omp_set_num_threads(g_threads);
#endif
}
int jpo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

#define NCELLS(g) ((int64_t)(g)->n[0] * (g)->n[1] * ((g)->ndim == 3 ? (g)->n[2] : 1))

/* ---- spacing accessors: particles.di.{vertex,center,velocity} -------------
 * range grids: scalar x[2]-x[1]   (src/Particles/particles_utils.jl:137-140)
 * array grids: diff(x)[i]         (src/Particles/particles_utils.jl:76-79)
 * @dxi / getindex_dxi             (src/Utils.jl:66-88)                      */
static inline double d_of(const double *x, int uniform, int i0) {
    return uniform ? x[1] - x[0] : x[i0 + 1] - x[i0];
}

/* ---- find_parent_cell_bisection (src/Utils.jl:117-130) -------------------
 * 1-based arithmetic kept literally (div(hi + seed, 2)).  len = length(x).   */
static inline int bisect1(double px, const double *x, int len, int seed1) {
    int lo = 1, hi = len, seed = seed1;
    for (;;) {
        if (x[seed - 1] <= px && px <= x[seed]) return seed;
        if (x[seed - 1] < px) { lo = seed; seed = (hi + seed) / 2; }
        else                  { hi = seed; seed = (lo + seed) / 2; }
    }
}

/* ---- lerp (src/Interpolations/ndlerp.jl:11-15) ---------------------------- */
static inline double lerp1(double t, double v0, double v1) {
    return fma(t, v1, fma(-t, v0, v0));
}
static inline double lerp2(const double *v, const double *t) {
    return lerp1(t[1], lerp1(t[0], v[0], v[1]), lerp1(t[0], v[2], v[3]));
}
static inline double lerp3(const double *v, const double *t) {
    double a = lerp1(t[1], lerp1(t[0], v[0], v[1]), lerp1(t[0], v[2], v[3]));
    double b = lerp1(t[1], lerp1(t[0], v[4], v[5]), lerp1(t[0], v[6], v[7]));
    return lerp1(t[2], a, b);
}
double jpo_lerp(int ndim, const double *v, const double *t) {
    return ndim == 1 ? lerp1(t[0], v[0], v[1]) : ndim == 2 ? lerp2(v, t) : lerp3(v, t);
}

/* ---- integrator stages (src/Advection/Euler.jl:1-6, RK2.jl:1-38) ----------
 * @muladd a + s*c*dt*v  ->  muladd((s*c)*dt, v, a)   (MuladdMacro 0.2.4)     */
void jpo_first_stage(int scheme, double alpha, double dt, int n, const double *v, const double *p, double *out) {
    double c = scheme == 0 ? 1.0 * dt : (1.0 * alpha) * dt;
    for (int i = 0; i < n; i++) out[i] = fma(c, v[i], p[i]);
}
void jpo_second_stage(double alpha, double dt, int n, const double *v0, const double *v1, const double *p, double *out) {
    if (alpha == 0.5) {
        for (int i = 0; i < n; i++) out[i] = fma(1.0 * dt, v1[i], p[i]);
    } else {
        double b = 0.5 * (1.0 / alpha);      /* half * inv(alpha) */
        double a = 1.0 - b;                  /* one - half*inv(alpha) (exact either way: 0.5*x is exact) */
        for (int i = 0; i < n; i++) {
            double inner = fma(b, v1[i], a * v0[i]);
            out[i] = fma(1.0 * dt, inner, p[i]);
        }
    }
}

/* ---- MQS (src/Interpolations/MQS.jl:1-158) and LinP (src/Particles/Advection/advection_LinP.jl:96-391)
 * velocity reconstructions.  F: column-major array of component `comp` (0-based), sizes nF; idx1: 1-based
 * corrected cell indices on that component's grid; v: the 2^N corners; t: normalised coordinates.
 * Both are literal, including the reference's quirks: the 3-D MQS takes the outer stencil nodes of its
 * "top"/"back" face from the SAME k (j) plane as the "bottom"/"front" one, the z-component's MQS
 * correction runs along x, and LinP's 3-D z-component corner tuple is ordered as the code orders it. */
static inline double Fat(const double *F, const int *nF, int N, int i1, int j1, int k1) {   /* 1-based */
    return F[(i1 - 1) + (int64_t)nF[0] * ((j1 - 1) + (N == 3 ? (int64_t)nF[1] * (k1 - 1) : 0))];
}
/* one 4-corner MQS sweep: v = (a0, a1 | b0, b1), quadratic correction along the first pair direction.
 * lo/hi outer nodes of the two edges are passed in (already fetched from F). */
static inline double mqs_edge(double v0e, double v1e, double tq, double outer_lo, double outer_hi) {
    const double half = 0.5;
    const double l = lerp1(tq, v0e, v1e);
    double a, b, c;
    if (tq < half) { a = outer_lo; b = v0e; c = v1e; } else { a = v0e; b = v1e; c = outer_hi; }
    const double corr = (half * ((tq - half) * (tq - half))) * (fma(-2.0, b, a) + c);
    return l + corr;
}
static double mqs_eval(const double *F, const int *nF, int N, int comp, const int *idx1, const double *v, const double *t) {
    const int i = idx1[0], j = idx1[1], k = N == 3 ? idx1[2] : 1;
    if (N == 2) {
        if (comp == 0) {
            double e0 = mqs_edge(v[0], v[1], t[0], Fat(F, nF, N, i - 1, j, 1), Fat(F, nF, N, i + 2, j, 1));
            double e1 = mqs_edge(v[2], v[3], t[0], Fat(F, nF, N, i - 1, j + 1, 1), Fat(F, nF, N, i + 2, j + 1, 1));
            return lerp1(t[1], e0, e1);
        }
        double e0 = mqs_edge(v[0], v[2], t[1], Fat(F, nF, N, i, j - 1, 1), Fat(F, nF, N, i, j + 2, 1));
        double e1 = mqs_edge(v[1], v[3], t[1], Fat(F, nF, N, i + 1, j - 1, 1), Fat(F, nF, N, i + 1, j + 2, 1));
        return lerp1(t[0], e0, e1);
    }
    if (comp == 0) {            /* MQS-x on v[1:4] and v[5:8], both with plane k; lerp in t3 */
        double f[2];
        for (int h = 0; h < 2; h++) {
            const double *w = v + 4 * h;
            double e0 = mqs_edge(w[0], w[1], t[0], Fat(F, nF, N, i - 1, j, k), Fat(F, nF, N, i + 2, j, k));
            double e1 = mqs_edge(w[2], w[3], t[0], Fat(F, nF, N, i - 1, j + 1, k), Fat(F, nF, N, i + 2, j + 1, k));
            f[h] = lerp1(t[1], e0, e1);
        }
        return lerp1(t[2], f[0], f[1]);
    }
    if (comp == 1) {            /* MQS-y on v[1:4] and v[5:8], both with plane k */
        double f[2];
        for (int h = 0; h < 2; h++) {
            const double *w = v + 4 * h;
            double e0 = mqs_edge(w[0], w[2], t[1], Fat(F, nF, N, i, j - 1, k), Fat(F, nF, N, i, j + 2, k));
            double e1 = mqs_edge(w[1], w[3], t[1], Fat(F, nF, N, i + 1, j - 1, k), Fat(F, nF, N, i + 1, j + 2, k));
            f[h] = lerp1(t[0], e0, e1);
        }
        return lerp1(t[2], f[0], f[1]);
    }
    /* MQS-z: front = (v1,v2,v5,v6), back = (v3,v4,v7,v8), each with (t1,t3) and the SAME j; correction along x */
    double f[2];
    for (int h = 0; h < 2; h++) {
        const double w[4] = {v[2 * h], v[2 * h + 1], v[4 + 2 * h], v[4 + 2 * h + 1]};
        double e0 = mqs_edge(w[0], w[1], t[0], Fat(F, nF, N, i - 1, j, k), Fat(F, nF, N, i + 2, j, k));
        double e1 = mqs_edge(w[2], w[3], t[0], Fat(F, nF, N, i - 1, j, k + 1), Fat(F, nF, N, i + 2, j, k + 1));
        f[h] = lerp1(t[2], e0, e1);
    }
    return lerp1(t[1], f[0], f[1]);
}

static inline int clampi(int x, int lo, int hi) { return x > hi ? hi : (x < lo ? lo : x); }
static double linp_eval(const double *F, const int *nF, int N, int comp, const int *idx1, const double *xc /* cell corner coords */,
                        const double *dxi, const double *p, double VL) {
    /* augment_offset(Val(comp+1)): rows of three offsets per direction */
    static const int T3[3] = {-1, 0, 1};
    int offI[4][3], offJ[4][3], offK[4][3];
    for (int r = 0; r < 4; r++)
        for (int m = 0; m < 3; m++) {
            const int bj = (r & 1), bk = (r >> 1);           /* rows: (0,0), (1,0), (0,1), (1,1) pattern of the tables */
            if (comp == 0)      { offI[r][m] = T3[m]; offJ[r][m] = bj;    offK[r][m] = bk; }
            else if (comp == 1) { offI[r][m] = bj;    offJ[r][m] = T3[m]; offK[r][m] = bk; }
            else                { offI[r][m] = bk;    offJ[r][m] = bj;    offK[r][m] = T3[m]; }
        }
    int i = idx1[0], j = idx1[1], k = N == 3 ? idx1[2] : 1;
    if (comp == 0) i += p[0] > xc[0] + dxi[0] / 2;
    if (comp == 1) j += p[1] > xc[1] + dxi[1] / 2;
    if (comp == 2) k += p[2] > xc[2] + dxi[2] / 2;
    /* rows used: 2-D (1,1),(2,2); 3-D (1,1,1),(2,2,2),(3,1,3),(4,2,4) (1-based rows of offset_i, offset_j, offset_k) */
    static const int RI[4] = {0, 1, 2, 3}, RJ[4] = {0, 1, 0, 1}, RK[4] = {0, 1, 2, 3};
    const int nrow = N == 2 ? 2 : 4;
    double av[8];
    for (int r = 0; r < nrow; r++) {
        double f[3];
        for (int m = 0; m < 3; m++)
            f[m] = Fat(F, nF, N, clampi(i + offI[RI[r]][m], 1, nF[0]), clampi(j + offJ[RJ[r]][m], 1, nF[1]),
                       N == 3 ? clampi(k + offK[RK[r]][m], 1, nF[2]) : 1);
        av[2 * r] = (f[0] + f[1]) / 2;
        av[2 * r + 1] = (f[2] + f[1]) / 2;
    }
    double FP[8];
    if (comp == 0) { for (int q = 0; q < 2 * nrow; q++) FP[q] = av[q]; }
    else if (N == 2) { FP[0] = av[0]; FP[1] = av[2]; FP[2] = av[1]; FP[3] = av[3]; }
    else { FP[0] = av[0]; FP[1] = av[2]; FP[2] = av[1]; FP[3] = av[3]; FP[4] = av[4]; FP[5] = av[6]; FP[6] = av[5]; FP[7] = av[7]; }
    double xP[3], tP[3];
    for (int d = 0; d < N; d++) xP[d] = xc[d];
    const int off = 1 - 2 * (p[comp] < xc[comp] + dxi[comp] / 2);
    xP[comp] = xc[comp] + ((double)off * dxi[comp]) / 2;
    for (int d = 0; d < N; d++) tP[d] = (p[d] - xP[d]) * (1.0 / dxi[d]);
    const double VP = N == 2 ? lerp2(FP, tP) : lerp3(FP, tP);
    const double A = 2.0 / 3.0;
    return A * VL + (1 - A) * VP;
}

/* ---- interp_velocity2particle (src/Particles/Advection/advection.jl:93-148,
 *      src/Advection/advection.jl:3-47, src/Interpolations/utils.jl:54-58) -- */
static int g_interp = 0;      /* 0 linear (advection!), 1 LinP (advection_LinP!), 2 MQS (advection_MQS!) */
static inline void interp_velocity(const jpo_grid *g, const double *const *V, const double *p,
                                   const int *cell1 /* 1-based storage cell */, double *vout) {
    const int N = g->ndim;
    for (int c = 0; c < N; c++) {
        /* check_local_limits: inclusive, all dims, this component's grid */
        int ok = 1;
        for (int d = 0; d < N; d++) {
            const double *x = g->xvel[c][d];
            if (!(x[0] <= p[d] && p[d] <= x[g->nvel[c][d] - 1])) { ok = 0; break; }
        }
        if (!ok) { vout[c] = INFINITY; continue; }
        int idx[3] = {1, 1, 1};
        double t[3];
        for (int d = 0; d < N; d++) {
            const double *x = g->xvel[c][d];
            idx[d] = bisect1(p[d], x, g->nvel[c][d], cell1[d]);
            double dx = d_of(x, g->uniform, idx[d] - 1);
            t[d] = (p[d] - x[idx[d] - 1]) * (1.0 / dx);
        }
        const double *F = V[c];
        const int64_t s1 = g->nvel[c][0], s2 = (int64_t)g->nvel[c][0] * g->nvel[c][1];
        const int64_t b = (idx[0] - 1) + s1 * (idx[1] - 1) + (N == 3 ? s2 * (idx[2] - 1) : 0);
        double v[8] = {F[b], F[b + 1], F[b + s1], F[b + s1 + 1], 0, 0, 0, 0};
        if (N == 3) { v[4] = F[b + s2]; v[5] = F[b + s2 + 1]; v[6] = F[b + s2 + s1]; v[7] = F[b + s2 + s1 + 1]; }
        const double VL = N == 2 ? lerp2(v, t) : lerp3(v, t);
        vout[c] = VL;
        if (g_interp == 0) continue;
        /* interior test: all(1 .< indices .< size(F) .- 1)  (advection_LinP.jl:113, advection_MQS.jl:117) */
        int interior = 1;
        for (int d = 0; d < N; d++) interior &= (1 < idx[d]) & (idx[d] < g->nvel[c][d] - 1);
        if (!interior) continue;
        if (g_interp == 2) { vout[c] = mqs_eval(F, g->nvel[c], N, c, idx, v, t); continue; }
        double xcn[3], dxi[3];
        for (int d = 0; d < N; d++) { xcn[d] = g->xvel[c][d][idx[d] - 1]; dxi[d] = d_of(g->xvel[c][d], g->uniform, idx[d] - 1); }
        vout[c] = linp_eval(F, g->nvel[c], N, c, idx, xcn, dxi, p, VL);
    }
}
/* single-point entry for the tests: velocity at p seeded from the 1-based cell cell1 */
void jpo_interp_velocity(const jpo_grid *g, const double *const *V, const double *p, const int *cell1, int interp, double *vout) {
    const int save = g_interp;
    g_interp = interp;
    interp_velocity(g, V, p, cell1, vout);
    g_interp = save;
}

/* ---- advect_particle (src/Particles/Advection/Euler.jl:1-20, RK2.jl:1-26,
 *      RK4.jl:1-19) -------------------------------------------------------- */
static inline void advect_particle(const jpo_grid *g, int scheme, double alpha, const double *const *V,
                                   double dt, const int *cell1, const double *p0, double *pout) {
    const int N = g->ndim;
    double k1[3], k2[3], k3[3], k4[3], q[3];
    interp_velocity(g, V, p0, cell1, k1);
    if (scheme == 0) { jpo_first_stage(0, alpha, dt, N, k1, p0, pout); return; }
    if (scheme == 1) {
        jpo_first_stage(1, alpha, dt, N, k1, p0, q);
        interp_velocity(g, V, q, cell1, k2);
        jpo_second_stage(alpha, dt, N, k1, k2, p0, pout);
        return;
    }
    /* RK4: unfused, left to right */
    for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k1[d] / 2;
    interp_velocity(g, V, q, cell1, k2);
    for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k2[d] / 2;
    interp_velocity(g, V, q, cell1, k3);
    for (int d = 0; d < N; d++) q[d] = p0[d] + dt * k3[d];
    interp_velocity(g, V, q, cell1, k4);
    for (int d = 0; d < N; d++)
        pout[d] = p0[d] + dt * (((k1[d] + 2 * k2[d]) + 2 * k3[d]) + k4[d]) / 6;
}

/* advection! (src/Particles/Advection/advection.jl:35-91) */
int jpo_advect(const jpo_grid *g, double *const *coords, const uint8_t *index, int scheme, double alpha,
               const double *const *V, double dt);
/* advection_LinP! / advection_MQS! (advection_LinP.jl:12-91, advection_MQS.jl:16-91): same drivers, other interpolant */
int jpo_advect_interp(const jpo_grid *g, double *const *coords, const uint8_t *index, int scheme, double alpha,
                      const double *const *V, double dt, int interp) {
    if (interp < 0 || interp > 2) return -1;
    g_interp = interp;
    const int rc = jpo_advect(g, coords, index, scheme, alpha, V, dt);
    g_interp = 0;
    return rc;
}
int jpo_advect(const jpo_grid *g, double *const *coords, const uint8_t *index, int scheme, double alpha,
               const double *const *V, double dt) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    if (scheme == 1 && !(0 < alpha && alpha < 1)) return -1;
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c = 0; c < C; c++) {
        int cell1[3] = {(int)(c % nx) + 1, (int)((c / nx) % ny) + 1, (int)(c / ((int64_t)nx * ny)) + 1};
        for (int s = 0; s < g->S; s++) {
            const int64_t e = c + (int64_t)s * C;
            if (!index[e]) continue;
            double p0[3], p1[3];
            for (int d = 0; d < N; d++) p0[d] = coords[d][e];
            advect_particle(g, scheme, alpha, V, dt, cell1, p0, p1);
            for (int d = 0; d < N; d++) coords[d][e] = p1[d];
        }
    }
    return 0;
}

/* ---- Philox4x32-10 (Salmon et al. 2011), counter-based RNG ---------------- */
static inline void philox4x32_10(uint32_t ctr[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static inline double u01(uint32_t hi, uint32_t lo) {   /* 53-bit uniform in [0,1) */
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * 0x1.0p-53;
}
/* three uniforms for (seed, purpose, step, cell, slot): ctr = {cell, slot, purpose, step<<1|sub} */
void jpo_rand3(uint64_t seed, uint32_t purpose, uint32_t step, uint32_t cell, uint32_t slot, double *r) {
    uint32_t a[4] = {cell, slot, purpose, (step << 1) | 0u};
    uint32_t b[4] = {cell, slot, purpose, (step << 1) | 1u};
    philox4x32_10(a, (uint32_t)seed, (uint32_t)(seed >> 32));
    philox4x32_10(b, (uint32_t)seed, (uint32_t)(seed >> 32));
    r[0] = u01(a[0], a[1]); r[1] = u01(a[2], a[3]); r[2] = u01(b[0], b[1]);
}

/* ---- init_particles / fill_coords_index! (src/Particles/particles_utils.jl:
 *      108-194, quadrant_masks :224-240) ----------------------------------- */
int jpo_init_particles(const jpo_grid *g, double *const *coords, uint8_t *index, int nxcell, uint64_t seed) {
    const int N = g->ndim, NQ = N == 2 ? 4 : 8;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    const int npq = (nxcell + NQ - 1) / NQ;
    if (npq * NQ > g->S) return -1;
    for (int d = 0; d < N; d++)
        for (int64_t e = 0; e < C * g->S; e++) coords[d][e] = NAN;
    memset(index, 0, (size_t)(C * g->S));
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c = 0; c < C; c++) {
        int ci[3] = {(int)(c % nx), (int)((c / nx) % ny), (int)(c / ((int64_t)nx * ny))};
        double x0[3], dx[3];
        for (int d = 0; d < N; d++) { x0[d] = g->xv[d][ci[d]]; dx[d] = d_of(g->xv[d], g->uniform, ci[d]); }
        int l = 0;
        for (int iq = 0; iq < NQ; iq++) {
            double xq[3];
            for (int d = 0; d < N; d++) xq[d] = x0[d] + dx[d] * (double)((iq >> d) & 1) / 2;
            for (int k = 0; k < npq; k++, l++) {
                double r[3];
                jpo_rand3(seed, 0u, 0u, (uint32_t)c, (uint32_t)l, r);
                const int64_t e = c + (int64_t)l * C;
                for (int d = 0; d < N; d++) coords[d][e] = xq[d] + dx[d] / 2 * r[d];
                index[e] = 1;
            }
        }
    }
    return 0;
}

/* ---- move_particles! (src/Particles/move_safe.jl:21-125; isincell
 *      src/Particles/utils.jl:7-15; indomain :199-206; find_free_memory
 *      :192-197) ------------------------------------------------------------ */
/* Slot policy.  0 = the reference's: the free-slot search starts at `starting_point`, which is carried
 * over from one migrant of the source cell to the next even when they go to DIFFERENT destination cells
 * (move_safe.jl:114-118).  1 = "compact" (an OPTION of the library, not reference behaviour): the search
 * starts at slot 0 for every migrant, i.e. the line `starting_point = free_idx` is dropped.  Everything
 * else (sweep order, strict tests, drop when no slot is found) is unchanged.
 * 2 = "dense" (also an option of the library): vacate everything, then place -- jpo_move_dense below. */
static int g_move_compact = 0, g_move_dense = 0;
void jpo_set_move_policy(int policy) { g_move_compact = policy == 1; g_move_dense = policy == 2; }

static void move_cell(const jpo_grid *g, double *const *coords, uint8_t *index, double *const *args, int nargs,
                      const int *ci /*0-based*/, int64_t *n_moved, int64_t *n_dropped, int64_t *n_deleted) {
    const int N = g->ndim, S = g->S;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    const int64_t c = ci[0] + (int64_t)nx * (ci[1] + (int64_t)ny * (N == 3 ? ci[2] : 0));
    double corner[3], dx[3], lo[3], hi[3];
    for (int d = 0; d < N; d++) {
        corner[d] = g->xv[d][ci[d]];
        dx[d] = d_of(g->xv[d], g->uniform, ci[d]);
        lo[d] = g->xv[d][0]; hi[d] = g->xv[d][g->n[d]];
    }
    int starting_point = 0;                         /* 0-based cursor */
    double cache[64];
    for (int ip = 0; ip < S; ip++) {
        const int64_t e = c + (int64_t)ip * C;
        if (!index[e]) continue;
        double p[3];
        int incell = 1, indom = 1;
        for (int d = 0; d < N; d++) {
            p[d] = coords[d][e];
            incell &= (corner[d] < p[d]) & (p[d] < corner[d] + dx[d]);
        }
        if (incell) continue;
        for (int d = 0; d < N; d++) if (!(lo[d] < p[d] && p[d] < hi[d])) { indom = 0; break; }
        if (!indom) {
            index[e] = 0;
            for (int d = 0; d < N; d++) coords[d][e] = NAN;
            for (int a = 0; a < nargs; a++) args[a][e] = NAN;
            (*n_deleted)++;
            continue;
        }
        int nc[3] = {0, 0, 0};
        for (int d = 0; d < N; d++) nc[d] = bisect1(p[d], g->xv[d], g->n[d] + 1, ci[d] + 1) - 1;
        for (int a = 0; a < nargs; a++) cache[a] = args[a][e];
        index[e] = 0;
        for (int d = 0; d < N; d++) coords[d][e] = NAN;
        for (int a = 0; a < nargs; a++) args[a][e] = NAN;
        const int64_t c2 = nc[0] + (int64_t)nx * (nc[1] + (int64_t)ny * (N == 3 ? nc[2] : 0));
        int free_idx = -1;
        for (int i = starting_point; i < S; i++)
            if (!index[c2 + (int64_t)i * C]) { free_idx = i; break; }
        if (free_idx < 0) { (*n_dropped)++; continue; }
        if (!g_move_compact) starting_point = free_idx;
        const int64_t e2 = c2 + (int64_t)free_idx * C;
        index[e2] = 1;
        for (int d = 0; d < N; d++) coords[d][e2] = p[d];
        for (int a = 0; a < nargs; a++) args[a][e2] = cache[a];
        (*n_moved)++;
    }
}

/* The "dense" policy (library option, not in the reference).  Pass 0: every live particle that fails the strict isincell test
 * of its cell leaves: its slot is vacated (mask 0, NaN) and its payload kept aside.  Pass 1: the leavers are visited in the
 * reference's order (3^N colours; within a colour the source cells never share a destination; slot order within a cell); one
 * outside the domain is deleted, any other goes to the LOWEST free slot of its destination cell (same bisection as the
 * reference) or is dropped when there is none. */
static int jpo_move_dense(const jpo_grid *g, double *const *coords, uint8_t *index, double *const *args, int nargs, int64_t *stats) {
    const int N = g->ndim, S = g->S, NA = N + nargs;
    const int64_t C = NCELLS(g), E = C * S;
    const int nx = g->n[0], ny = g->n[1];
    double *arr[3 + 64];
    for (int d = 0; d < N; d++) arr[d] = coords[d];
    for (int a = 0; a < nargs; a++) arr[N + a] = args[a];
    uint8_t *leaves = (uint8_t *)calloc((size_t)E, 1);
    double *keep = (double *)malloc((size_t)E * NA * sizeof(double));      /* payload of the leavers, by element */
    if (!leaves || !keep) { free(leaves); free(keep); return -2; }
    for (int64_t c = 0; c < C; c++) {
        const int ci[3] = {(int)(c % nx), (int)((c / nx) % ny), (int)(c / ((int64_t)nx * ny))};
        for (int ip = 0; ip < S; ip++) {
            const int64_t e = c + (int64_t)ip * C;
            if (!index[e]) continue;
            int incell = 1;
            for (int d = 0; d < N; d++) {
                const double corner = g->xv[d][ci[d]], p = coords[d][e];
                incell &= (corner < p) & (p < corner + d_of(g->xv[d], g->uniform, ci[d]));
            }
            if (incell) continue;
            leaves[e] = 1;
            for (int a = 0; a < NA; a++) { keep[(size_t)e * NA + a] = arr[a][e]; arr[a][e] = NAN; }
            index[e] = 0;
        }
    }
    int64_t moved = 0, dropped = 0, deleted = 0;
    const int ncol[3] = {(nx + 2) / 3, (ny + 2) / 3, N == 3 ? (g->n[2] + 2) / 3 : 1};
    for (int ox = 0; ox < 3; ox++)
        for (int oy = 0; oy < 3; oy++)
            for (int oz = 0; oz < (N == 3 ? 3 : 1); oz++)
                for (int K = 0; K < ncol[2]; K++)
                    for (int J = 0; J < ncol[1]; J++)
                        for (int I = 0; I < ncol[0]; I++) {
                            const int ci[3] = {3 * I + ox, 3 * J + oy, N == 3 ? 3 * K + oz : 0};
                            if (ci[0] >= nx || ci[1] >= ny || (N == 3 && ci[2] >= g->n[2])) continue;
                            const int64_t c = ci[0] + (int64_t)nx * (ci[1] + (int64_t)ny * ci[2]);
                            for (int ip = 0; ip < S; ip++) {
                                const int64_t e = c + (int64_t)ip * C;
                                if (!leaves[e]) continue;
                                const double *p = keep + (size_t)e * NA;
                                int indom = 1, nc[3] = {0, 0, 0};
                                for (int d = 0; d < N; d++) if (!(g->xv[d][0] < p[d] && p[d] < g->xv[d][g->n[d]])) { indom = 0; break; }
                                if (!indom) { deleted++; continue; }
                                for (int d = 0; d < N; d++) nc[d] = bisect1(p[d], g->xv[d], g->n[d] + 1, ci[d] + 1) - 1;
                                const int64_t c2 = nc[0] + (int64_t)nx * (nc[1] + (int64_t)ny * nc[2]);
                                int free_idx = -1;
                                for (int i = 0; i < S; i++) if (!index[c2 + (int64_t)i * C]) { free_idx = i; break; }
                                if (free_idx < 0) { dropped++; continue; }
                                const int64_t e2 = c2 + (int64_t)free_idx * C;
                                index[e2] = 1;
                                for (int a = 0; a < NA; a++) arr[a][e2] = p[a];
                                moved++;
                            }
                        }
    free(leaves); free(keep);
    if (stats) { stats[0] = moved; stats[1] = dropped; stats[2] = deleted; }
    return 0;
}

int jpo_move(const jpo_grid *g, double *const *coords, uint8_t *index, double *const *args, int nargs,
             int64_t *stats /* moved, dropped, deleted */) {
    const int N = g->ndim;
    if (nargs > 64) return -1;
    if (g_move_dense) return jpo_move_dense(g, coords, index, args, nargs, stats);
    int ncol[3] = {(g->n[0] + 2) / 3, (g->n[1] + 2) / 3, N == 3 ? (g->n[2] + 2) / 3 : 1};
    int64_t moved = 0, dropped = 0, deleted = 0;
    const int oz_max = N == 3 ? 3 : 1;
    for (int ox = 0; ox < 3; ox++)
        for (int oy = 0; oy < 3; oy++)
            for (int oz = 0; oz < oz_max; oz++) {
                const int64_t nt = (int64_t)ncol[0] * ncol[1] * ncol[2];
#pragma omp parallel for schedule(static) reduction(+ : moved, dropped, deleted) if (g_threads > 1)
                for (int64_t t = 0; t < nt; t++) {
                    int I = (int)(t % ncol[0]), J = (int)((t / ncol[0]) % ncol[1]), K = (int)(t / ((int64_t)ncol[0] * ncol[1]));
                    int ci[3] = {3 * I + ox, 3 * J + oy, N == 3 ? 3 * K + oz : 0};
                    if (ci[0] >= g->n[0] || ci[1] >= g->n[1] || (N == 3 && ci[2] >= g->n[2])) continue;
                    move_cell(g, coords, index, args, nargs, ci, &moved, &dropped, &deleted);
                }
            }
    if (stats) { stats[0] = moved; stats[1] = dropped; stats[2] = deleted; }
    return 0;
}

/* ---- force_injection! (src/Particles/forced_injection.jl:16-79) -------------
 * pnew[d][c + k*C] = component d of p_new[I..., k]; a cell injects iff its first entry is not NaN (:36, :60; the
 * reference tests overload isnan for their point type, test/test_2D.jl:81); the helper counter `c` of :37-40 advances
 * with the slot loop, so free slot ip always takes entry ip. */
int jpo_force_injection(const jpo_grid *g, double *const *coords, uint8_t *index, const double *const *pnew, double *const *fields,
                        const double *values, int nfields) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    for (int64_t cell = 0; cell < C; cell++) {
        if (isnan(pnew[0][cell])) continue;
        int c = 0;
        for (int ip = 0; ip < g->S; ip++) {
            c += 1;
            if (c > g->S) continue;
            const int64_t e = cell + (int64_t)ip * C;
            if (index[e]) continue;
            for (int d = 0; d < N; d++) coords[d][e] = pnew[d][cell + (int64_t)(c - 1) * C];
            index[e] = 1;
            for (int a = 0; a < nfields; a++) fields[a][e] = values[a];
        }
    }
    return 0;
}

/* ---- clean_particles! (src/Particles/move_safe.jl:289-320) ---------------- */
int jpo_clean(const jpo_grid *g, double *const *coords, uint8_t *index, double *const *args, int nargs) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c = 0; c < C; c++) {
        int ci[3] = {(int)(c % nx), (int)((c / nx) % ny), (int)(c / ((int64_t)nx * ny))};
        for (int s = 0; s < g->S; s++) {
            const int64_t e = c + (int64_t)s * C;
            if (!index[e]) continue;
            int incell = 1;
            for (int d = 0; d < N; d++) {
                double xvd = g->xv[d][ci[d]], dx = d_of(g->xv[d], g->uniform, ci[d]), p = coords[d][e];
                incell &= (xvd < p) & (p < xvd + dx);
            }
            if (incell) continue;
            index[e] = 0;
            for (int d = 0; d < N; d++) coords[d][e] = NAN;
            for (int a = 0; a < nargs; a++) args[a][e] = NAN;
        }
    }
    return 0;
}

/* ---- inject_particles! (src/Particles/injection.jl:19-131; index_min_distance
 *      :330-393; new_particle :411-417; quadrant_corners :442-461; distance
 *      src/Interpolations/utils.jl:9-19) ------------------------------------ */
static void inject_cell(const jpo_grid *g, double *const *coords, uint8_t *index, double *const *args, int nargs,
                        int min_xcell, uint64_t seed, uint32_t step, const int *ci, int64_t *n_injected) {
    const int N = g->ndim, S = g->S, NQ = N == 2 ? 4 : 8;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1], nz = N == 3 ? g->n[2] : 1;
    const int64_t c = ci[0] + (int64_t)nx * (ci[1] + (int64_t)ny * (N == 3 ? ci[2] : 0));
    double xvc[3], dq[3];
    for (int d = 0; d < N; d++) { xvc[d] = g->xv[d][ci[d]]; dq[d] = d_of(g->xv[d], g->uniform, ci[d]) / 2; }
    const int min_xq = (min_xcell + NQ - 1) / NQ;      /* cld(min_xcell, NQ) */
    for (int iq = 0; iq < NQ; iq++) {
        double vq[3];
        for (int d = 0; d < N; d++) vq[d] = xvc[d] + dq[d] * (double)((iq >> d) & 1);
        int num = 0;
        for (int i = 0; i < S; i++) {
            const int64_t e = c + (int64_t)i * C;
            if (!index[e]) continue;
            int in = 1;
            for (int d = 0; d < N; d++) { double p = coords[d][e]; in &= (vq[d] < p) & (p < vq[d] + dq[d]); }
            num += in;
        }
        if (num >= min_xq) break;                      /* leaves the quadrant loop (injection.jl:100) */
        for (int i = 0; i < S; i++) {
            const int64_t e = c + (int64_t)i * C;
            if (index[e]) continue;
            num++;
            double r[3], pn[3];
            jpo_rand3(seed, 1u, step, (uint32_t)c, (uint32_t)i, r);
            for (int d = 0; d < N; d++) pn[d] = vq[d] + dq[d] * fma(0.95, r[d], 0.05);
            for (int d = 0; d < N; d++) coords[d][e] = pn[d];
            index[e] = 1;
            (*n_injected)++;
            /* nearest live particle in the 3^N neighbourhood, k,j,i outer, slot inner, strict < */
            double dmin = INFINITY; int64_t emin = -1;
            for (int kk = (N == 3 ? ci[2] - 1 : 0); kk <= (N == 3 ? ci[2] + 1 : 0); kk++)
                for (int jj = ci[1] - 1; jj <= ci[1] + 1; jj++)
                    for (int ii = ci[0] - 1; ii <= ci[0] + 1; ii++) {
                        if (ii < 0 || jj < 0 || kk < 0 || ii >= nx || jj >= ny || kk >= nz) continue;
                        const int64_t c2 = ii + (int64_t)nx * (jj + (int64_t)ny * kk);
                        for (int ip = 0; ip < S; ip++) {
                            if (c2 == c && ip == i) continue;
                            const int64_t e2 = c2 + (int64_t)ip * C;
                            if (!index[e2]) continue;
                            double s = 0;
                            for (int d = 0; d < N; d++) {
                                double del = coords[d][e2] - pn[d];
                                s = d == 0 ? del * del : s + del * del;
                            }
                            double dist = sqrt(s);
                            if (dist < dmin) { dmin = dist; emin = e2; }
                        }
                    }
            if (emin >= 0)
                for (int a = 0; a < nargs; a++) args[a][e] = args[a][emin];
            /* no live donor anywhere: reference reads out of bounds (UB); args left untouched here */
            if (num >= min_xq) break;
        }
    }
}

int jpo_inject(const jpo_grid *g, double *const *coords, uint8_t *index, double *const *args, int nargs,
               int min_xcell, uint64_t seed, uint32_t step, int64_t *n_injected_out) {
    const int N = g->ndim;
    int ncol[3] = {(g->n[0] + 1) / 2, (g->n[1] + 1) / 2, N == 3 ? (g->n[2] + 1) / 2 : 1};
    int64_t injected = 0;
    const int oz_max = N == 3 ? 2 : 1;
    for (int ox = 0; ox < 2; ox++)
        for (int oy = 0; oy < 2; oy++)
            for (int oz = 0; oz < oz_max; oz++) {
                const int64_t nt = (int64_t)ncol[0] * ncol[1] * ncol[2];
#pragma omp parallel for schedule(static) reduction(+ : injected) if (g_threads > 1)
                for (int64_t t = 0; t < nt; t++) {
                    int I = (int)(t % ncol[0]), J = (int)((t / ncol[0]) % ncol[1]), K = (int)(t / ((int64_t)ncol[0] * ncol[1]));
                    int ci[3] = {2 * I + ox, 2 * J + oy, N == 3 ? 2 * K + oz : 0};
                    if (ci[0] >= g->n[0] || ci[1] >= g->n[1] || (N == 3 && ci[2] >= g->n[2])) continue;
                    inject_cell(g, coords, index, args, nargs, min_xcell, seed, step, ci, &injected);
                }
            }
    if (n_injected_out) *n_injected_out = injected;
    return 0;
}

/* ---- inject_particles_phase! (src/Particles/injection.jl:146-325) -------------------------
 * Differences from inject_particles!: EVERY quadrant is examined (`continue`, :271, instead of the
 * `break` of :100); the nearest particle (index_min_distance, looked up BEFORE the new particle is
 * marked live, :283) only donates its PHASE; the other particle fields args[j] are interpolated
 * from grid fields[j] at the new position and clamped to the extrema of the interpolation stencil:
 * centre fields (size == cells, :295-305) from the cell-centre grid with the shifted / clamped cell
 * index of centroid2particle, everything else (:307-311) as a vertex field of the storage cell.
 * fkind[j]: 1 = centre field (n), 0 = vertex field (n+1).  RNG purpose 2. */
static inline double clamp_julia(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

static void inject_phase_cell(const jpo_grid *g, double *const *coords, uint8_t *index, double *phases, double *const *args,
                              const double *const *fields, const int *fkind, int nargs, int min_xcell, uint64_t seed,
                              uint32_t step, const int *ci, int64_t *n_injected) {
    const int N = g->ndim, S = g->S, NQ = N == 2 ? 4 : 8;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1], nz = N == 3 ? g->n[2] : 1;
    const int64_t c = ci[0] + (int64_t)nx * (ci[1] + (int64_t)ny * (N == 3 ? ci[2] : 0));
    double xvc[3], dcell[3], dq[3], xcc[3];
    for (int d = 0; d < N; d++) {
        xvc[d] = g->xv[d][ci[d]];
        dcell[d] = d_of(g->xv[d], g->uniform, ci[d]);
        dq[d] = dcell[d] / 2;
        xcc[d] = (xvc[d] + dq[d] * 0.0) + dq[d];          /* xvi_quadrants[1] .+ di_quadrant */
    }
    const int min_xq = (min_xcell + NQ - 1) / NQ;
    for (int iq = 0; iq < NQ; iq++) {
        double vq[3];
        for (int d = 0; d < N; d++) vq[d] = xvc[d] + dq[d] * (double)((iq >> d) & 1);
        int num = 0;
        for (int i = 0; i < S; i++) {
            const int64_t e = c + (int64_t)i * C;
            if (!index[e]) continue;
            int in = 1;
            for (int d = 0; d < N; d++) { double p = coords[d][e]; in &= (vq[d] < p) & (p < vq[d] + dq[d]); }
            num += in;
        }
        if (num >= min_xq) continue;
        for (int i = 0; i < S; i++) {
            const int64_t e = c + (int64_t)i * C;
            if (index[e]) continue;
            num++;
            double r[3], pn[3];
            jpo_rand3(seed, 2u, step, (uint32_t)c, (uint32_t)i, r);
            for (int d = 0; d < N; d++) pn[d] = vq[d] + dq[d] * fma(0.95, r[d], 0.05);
            /* phase of the nearest live particle in the 3^N neighbourhood (k,j,i outer, slot inner, strict <) */
            double dmin = INFINITY; int64_t emin = -1;
            for (int kk = (N == 3 ? ci[2] - 1 : 0); kk <= (N == 3 ? ci[2] + 1 : 0); kk++)
                for (int jj = ci[1] - 1; jj <= ci[1] + 1; jj++)
                    for (int ii = ci[0] - 1; ii <= ci[0] + 1; ii++) {
                        if (ii < 0 || jj < 0 || kk < 0 || ii >= nx || jj >= ny || kk >= nz) continue;
                        const int64_t c2 = ii + (int64_t)nx * (jj + (int64_t)ny * kk);
                        for (int ip = 0; ip < S; ip++) {
                            if (c2 == c && ip == i) continue;
                            const int64_t e2 = c2 + (int64_t)ip * C;
                            if (!index[e2]) continue;
                            double s = 0;
                            for (int d = 0; d < N; d++) {
                                double del = coords[d][e2] - pn[d];
                                s = d == 0 ? del * del : s + del * del;
                            }
                            double dist = sqrt(s);
                            if (dist < dmin) { dmin = dist; emin = e2; }
                        }
                    }
            if (emin >= 0) phases[e] = phases[emin];       /* no live donor: the reference indexes slot 0 (UB); left untouched */
            for (int d = 0; d < N; d++) coords[d][e] = pn[d];
            index[e] = 1;
            (*n_injected)++;
            for (int a = 0; a < nargs; a++) {
                const double *F = fields[a];
                int ic[3] = {0, 0, 0};
                double t[3], v[8];
                int64_t s1, s2;
                if (fkind[a] == 1) {
                    for (int d = 0; d < N; d++) {
                        int i1 = ci[d] + 1;                 /* 1-based */
                        if (pn[d] < xcc[d]) i1 -= 1;
                        const int hi1 = g->n[d] - 1;
                        if (i1 > hi1) i1 = hi1;             /* clamp(x, 1, sz-1): x > hi first */
                        else if (i1 < 1) i1 = 1;
                        ic[d] = i1 - 1;
                        t[d] = (pn[d] - g->xc[d][ic[d]]) * (1.0 / d_of(g->xc[d], g->uniform, ic[d]));
                    }
                    s1 = nx; s2 = (int64_t)nx * ny;
                } else {
                    for (int d = 0; d < N; d++) { ic[d] = ci[d]; t[d] = (pn[d] - g->xv[d][ci[d]]) * (1.0 / dcell[d]); }
                    s1 = nx + 1; s2 = (int64_t)(nx + 1) * (ny + 1);
                }
                const int64_t b = ic[0] + s1 * ic[1] + (N == 3 ? s2 * ic[2] : 0);
                v[0] = F[b]; v[1] = F[b + 1]; v[2] = F[b + s1]; v[3] = F[b + s1 + 1];
                if (N == 3) { v[4] = F[b + s2]; v[5] = F[b + s2 + 1]; v[6] = F[b + s2 + s1]; v[7] = F[b + s2 + s1 + 1]; }
                const double tmp = N == 2 ? lerp2(v, t) : lerp3(v, t);
                double lo = v[0], hi = v[0];
                for (int q = 1; q < (N == 2 ? 4 : 8); q++) { lo = v[q] < lo ? v[q] : lo; hi = v[q] > hi ? v[q] : hi; }
                args[a][e] = clamp_julia(tmp, lo, hi);
            }
            if (num >= min_xq) break;
        }
    }
}

int jpo_inject_phase(const jpo_grid *g, double *const *coords, uint8_t *index, double *phases, double *const *args,
                     const double *const *fields, const int *fkind, int nargs, int min_xcell, uint64_t seed, uint32_t step,
                     int64_t *n_injected_out) {
    const int N = g->ndim;
    int ncol[3] = {(g->n[0] + 1) / 2, (g->n[1] + 1) / 2, N == 3 ? (g->n[2] + 1) / 2 : 1};
    int64_t injected = 0;
    const int oz_max = N == 3 ? 2 : 1;
    for (int ox = 0; ox < 2; ox++)
        for (int oy = 0; oy < 2; oy++)
            for (int oz = 0; oz < oz_max; oz++) {
                const int64_t nt = (int64_t)ncol[0] * ncol[1] * ncol[2];
#pragma omp parallel for schedule(static) reduction(+ : injected) if (g_threads > 1)
                for (int64_t t = 0; t < nt; t++) {
                    int I = (int)(t % ncol[0]), J = (int)((t / ncol[0]) % ncol[1]), K = (int)(t / ((int64_t)ncol[0] * ncol[1]));
                    int ci[3] = {2 * I + ox, 2 * J + oy, N == 3 ? 2 * K + oz : 0};
                    if (ci[0] >= g->n[0] || ci[1] >= g->n[1] || (N == 3 && ci[2] >= g->n[2])) continue;
                    inject_phase_cell(g, coords, index, phases, args, fields, fkind, nargs, min_xcell, seed, step, ci, &injected);
                }
            }
    if (n_injected_out) *n_injected_out = injected;
    return 0;
}

/* ---- grid2particle! (src/Interpolations/grid_to_particle.jl:26-82, :265-273;
 *      field_corners src/Interpolations/utils.jl:98-118) -------------------- */
int jpo_grid2particle(const jpo_grid *g, const double *const *coords, const uint8_t *index, double *Fp, const double *F) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    const int64_t s1 = nx + 1, s2 = (int64_t)(nx + 1) * (ny + 1);
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c = 0; c < C; c++) {
        int ci[3] = {(int)(c % nx), (int)((c / nx) % ny), (int)(c / ((int64_t)nx * ny))};
        const int64_t b = ci[0] + s1 * ci[1] + (N == 3 ? s2 * ci[2] : 0);
        double v[8], xcn[3], idx[3];
        v[0] = F[b]; v[1] = F[b + 1]; v[2] = F[b + s1]; v[3] = F[b + s1 + 1];
        if (N == 3) { v[4] = F[b + s2]; v[5] = F[b + s2 + 1]; v[6] = F[b + s2 + s1]; v[7] = F[b + s2 + s1 + 1]; }
        for (int d = 0; d < N; d++) { xcn[d] = g->xv[d][ci[d]]; idx[d] = 1.0 / d_of(g->xv[d], g->uniform, ci[d]); }
        for (int s = 0; s < g->S; s++) {
            const int64_t e = c + (int64_t)s * C;
            if (!index[e]) continue;
            double t[3];
            for (int d = 0; d < N; d++) t[d] = (coords[d][e] - xcn[d]) * idx[d];
            Fp[e] = N == 2 ? lerp2(v, t) : lerp3(v, t);
        }
    }
    return 0;
}

/* ---- grid2particle_flip! (src/Interpolations/grid_to_particle.jl:125-173): PIC/FLIP blend.
 * di = grid_size(xvi) = abs(minimum(diff(x))) per dimension -- a SCALAR even on vector grids
 * (src/Interpolations/utils.jl:61-63); t = (p - xv[idx]) * inv(di);
 * Fp = muladd(F_pic, alpha, (Fp + (F_pic - F0_pic)) * (1 - alpha)). */
int jpo_grid2particle_flip(const jpo_grid *g, const double *const *coords, const uint8_t *index, double *Fp, const double *F,
                           const double *F0, double alpha) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    const int64_t s1 = nx + 1, s2 = (int64_t)(nx + 1) * (ny + 1);
    double idi[3];
    for (int d = 0; d < N; d++) {
        double m = g->xv[d][1] - g->xv[d][0];
        for (int i = 1; i < g->n[d]; i++) { double q = g->xv[d][i + 1] - g->xv[d][i]; m = q < m ? q : m; }
        idi[d] = 1.0 / fabs(m);
    }
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c = 0; c < C; c++) {
        int ci[3] = {(int)(c % nx), (int)((c / nx) % ny), (int)(c / ((int64_t)nx * ny))};
        const int64_t b = ci[0] + s1 * ci[1] + (N == 3 ? s2 * ci[2] : 0);
        double v[8], v0[8];
        for (int q = 0; q < (N == 2 ? 4 : 8); q++) {
            const int64_t o = b + (q & 1) + ((q >> 1) & 1) * s1 + ((q >> 2) & 1) * s2;
            v[q] = F[o]; v0[q] = F0[o];
        }
        for (int s = 0; s < g->S; s++) {
            const int64_t e = c + (int64_t)s * C;
            if (!index[e]) continue;
            double t[3];
            for (int d = 0; d < N; d++) t[d] = (coords[d][e] - g->xv[d][ci[d]]) * idi[d];
            const double Fpic = N == 2 ? lerp2(v, t) : lerp3(v, t);
            const double F0pic = N == 2 ? lerp2(v0, t) : lerp3(v0, t);
            const double Fflip = Fp[e] + (Fpic - F0pic);
            Fp[e] = fma(Fpic, alpha, Fflip * (1.0 - alpha));
        }
    }
    return 0;
}

/* ---- subgrid_diffusion! building blocks (src/Physics/subgrid_diffusion.jl:55-143).
 * The driver is a composition: memcopy (all slots) -> grid2particle! -> subgrid_diffusion_kernel! (live
 * slots) -> particle2grid! -> update_dT_subgrid_kernel! (grid) -> grid2particle! -> update_particle_
 * temperature_kernel! (all slots).  The three elementwise kernels are restated here; the drivers
 * (vertex and centroid variants) are composed in oracle.py from these and the interpolation kernels.
 * exp() is libm's: against CUDA's exp the stated 1e-12 tolerance applies (not bit-exactness). */
void jpo_subgrid_kernel(const jpo_grid *g, const uint8_t *index, const double *pT, double *pT0, double *pdT, const double *dt0,
                        double d, double dt) {
    const int64_t tot = NCELLS(g) * g->S;
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t e = 0; e < tot; e++) {
        if (!index[e]) continue;
        const double den = dt0[e] > 1.0e-9 ? dt0[e] : 1.0e-9;          /* max(dt0, dt_floor); NaN dt0 -> NaN as in Julia's max */
        const double denj = isnan(dt0[e]) ? dt0[e] : den;
        const double dTi = (pT[e] - pT0[e]) * (1 - exp(-d * dt / denj));
        pT0[e] = pT0[e] + dTi;
        pdT[e] = dTi;
    }
}
/* dTsubgrid[I] = dT[I + 1] - dTsubgrid[I]: dT carries one ghost node on the low side of every dimension */
void jpo_update_dT_subgrid(int N, const int *nsub /* extents of dTsubgrid */, const int *ndT /* extents of dT */, double *dTsub, const double *dT) {
    const int n2 = N == 3 ? nsub[2] : 1;
    for (int k = 0; k < n2; k++)
        for (int j = 0; j < nsub[1]; j++)
            for (int i = 0; i < nsub[0]; i++) {
                const int64_t a = i + (int64_t)nsub[0] * (j + (int64_t)nsub[1] * k);
                const int64_t b = (i + 1) + (int64_t)ndT[0] * ((j + 1) + (N == 3 ? (int64_t)ndT[1] * (k + 1) : 0));
                dTsub[a] = dT[b] - dTsub[a];
            }
}

/* ---- centroid2particle! (src/Interpolations/centroid_to_particle.jl:13-76) - */
int jpo_centroid2particle(const jpo_grid *g, const double *const *coords, double *Fp, const double *Fc) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    const int64_t s1 = nx, s2 = (int64_t)nx * ny;
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c = 0; c < C; c++) {
        int ci[3] = {(int)(c % nx), (int)((c / nx) % ny), (int)(c / ((int64_t)nx * ny))};
        for (int s = 0; s < g->S; s++) {
            const int64_t e = c + (int64_t)s * C;
            double p[3], t[3];
            int anynan = 0, cc[3] = {0, 0, 0};
            for (int d = 0; d < N; d++) { p[d] = coords[d][e]; anynan |= isnan(p[d]); }
            if (anynan) continue;
            for (int d = 0; d < N; d++) {
                int i1 = ci[d] + 1;                           /* 1-based */
                if (p[d] < g->xc[d][ci[d]]) i1 -= 1;           /* shifted_index */
                int hi1 = g->n[d] - 1;                        /* clamp(., 1, size(F)-1) */
                if (i1 > hi1) i1 = hi1;                       /* Base.clamp: x > hi ? hi : (x < lo ? lo : x) */
                else if (i1 < 1) i1 = 1;
                cc[d] = i1 - 1;
                double dx = d_of(g->xc[d], g->uniform, cc[d]);
                t[d] = (p[d] - g->xc[d][cc[d]]) * (1.0 / dx);
            }
            const int64_t b = cc[0] + s1 * cc[1] + (N == 3 ? s2 * cc[2] : 0);
            double v[8];
            v[0] = Fc[b]; v[1] = Fc[b + 1]; v[2] = Fc[b + s1]; v[3] = Fc[b + s1 + 1];
            if (N == 3) { v[4] = Fc[b + s2]; v[5] = Fc[b + s2 + 1]; v[6] = Fc[b + s2 + s1]; v[7] = Fc[b + s2 + s1 + 1]; }
            Fp[e] = N == 2 ? lerp2(v, t) : lerp3(v, t);
        }
    }
    return 0;
}

/* ---- particle2grid! (src/Interpolations/particle_to_grid.jl:23-68 (2D),
 *      :113-151 (3D), distance_weight :203-205) ----------------------------- */
int jpo_particle2grid(const jpo_grid *g, const double *const *coords, const uint8_t *index, double *F, const double *Fp) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1], nz = N == 3 ? g->n[2] : 0;
    const int64_t NN = (int64_t)(nx + 1) * (ny + 1) * (N == 3 ? nz + 1 : 1);
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t nd = 0; nd < NN; nd++) {
        int in = (int)(nd % (nx + 1)), jn = (int)((nd / (nx + 1)) % (ny + 1)), kn = (int)(nd / ((int64_t)(nx + 1) * (ny + 1)));
        double xn[3] = {g->xv[0][in], g->xv[1][jn], N == 3 ? g->xv[2][kn] : 0.0};
        double w = 0.0, wF = 0.0;
        for (int ko = (N == 3 ? -1 : 0); ko <= 0; ko++) {
            int kc = kn + ko;
            if (N == 3 && (kc < 0 || kc >= nz)) continue;
            for (int jo = -1; jo <= 0; jo++) {
                int jc = jn + jo;
                if (jc < 0 || jc >= ny) continue;
                for (int io = -1; io <= 0; io++) {
                    int ic = in + io;
                    if (ic < 0 || ic >= nx) continue;
                    const int64_t c = ic + (int64_t)nx * (jc + (int64_t)ny * (N == 3 ? kc : 0));
                    for (int s = 0; s < g->S; s++) {
                        const int64_t e = c + (int64_t)s * C;
                        if (!index[e]) continue;
                        double ss = 0;
                        for (int d = 0; d < N; d++) {
                            double del = xn[d] - coords[d][e];
                            ss = d == 0 ? del * del : ss + del * del;
                        }
                        double dist = sqrt(ss);
                        double wi = 1.0 / (dist * dist);
                        w += wi;
                        wF = fma(wi, Fp[e], wF);
                    }
                }
            }
        }
        F[nd] = N == 2 ? wF / w : wF * (1.0 / w);
    }
    return 0;
}

/* bilinear_weight (src/Interpolations/particle_to_grid.jl:211-223,
 * src/PhaseRatios/utils.jl:64-74): prod_d muladd(-|a-b|, inv(d), 1) */
static inline double bilinear_weight(int N, const double *a, const double *b, const double *di) {
    double val = 1.0;
    for (int d = 0; d < N; d++) val *= fma(-fabs(a[d] - b[d]), 1.0 / di[d], 1.0);
    return val;
}

/* ---- particle2centroid! (src/Interpolations/particle_to_grid_centroid.jl:
 *      10-45 (2D: any(isnan), w/F), :78-99 (3D: isnan(px), wF*inv(w))) ------ */
int jpo_particle2centroid(const jpo_grid *g, const double *const *coords, double *Fc, const double *Fp) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c = 0; c < C; c++) {
        int ci[3] = {(int)(c % nx), (int)((c / nx) % ny), (int)(c / ((int64_t)nx * ny))};
        double xcn[3], di[3];
        for (int d = 0; d < N; d++) { xcn[d] = g->xc[d][ci[d]]; di[d] = d_of(g->xv[d], g->uniform, ci[d]); }
        double w = 0.0, wF = 0.0;
        for (int s = 0; s < g->S; s++) {
            const int64_t e = c + (int64_t)s * C;
            double p[3];
            for (int d = 0; d < N; d++) p[d] = coords[d][e];
            if (N == 2 ? (isnan(p[0]) || isnan(p[1])) : isnan(p[0])) continue;
            double wi = bilinear_weight(N, xcn, p, di);
            w += wi;
            wF = fma(wi, Fp[e], wF);
        }
        Fc[c] = N == 2 ? wF / w : wF * (1.0 / w);
    }
    return 0;
}

/* ---- phase_ratios_center! (src/PhaseRatios/centers.jl:3-30;
 *      phase_ratio_weights src/PhaseRatios/utils.jl:45-62) ------------------
 * ratios layout: CellArray data[C, K] -> element (cell c, phase k) at c + k*C */
int jpo_phase_ratios_center(const jpo_grid *g, const double *const *coords, double *ratios, const double *phases, int K) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    if (K > 64) return -1;
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c = 0; c < C; c++) {
        int ci[3] = {(int)(c % nx), (int)((c / nx) % ny), (int)(c / ((int64_t)nx * ny))};
        double xcn[3], di[3], w[64];
        for (int d = 0; d < N; d++) { xcn[d] = g->xc[d][ci[d]]; di[d] = d_of(g->xv[d], g->uniform, ci[d]); }
        for (int k = 0; k < K; k++) w[k] = 0.0;
        for (int s = 0; s < g->S; s++) {
            const int64_t e = c + (int64_t)s * C;
            double p[3];
            for (int d = 0; d < N; d++) p[d] = coords[d][e];
            if (isnan(p[0])) continue;
            double x = bilinear_weight(N, xcn, p, di);
            double ph = phases[e];
            /* w .+ x .* (ph == j): x*false is a strong zero (copysign(0,x)) */
            for (int k = 0; k < K; k++) w[k] = w[k] + (ph == (double)(k + 1) ? x : copysign(0.0, x));
        }
        double sum = w[0];
        for (int k = 1; k < K; k++) sum = sum + w[k];
        double inv = 1.0 / sum;
        for (int k = 0; k < K; k++) ratios[c + (int64_t)k * C] = w[k] * inv;
    }
    return 0;
}

/* accumulate_weight / phase_ratio_weights inner step: w .+ x .* (phase == j), x*false = strong zero */
static inline void acc_phase(double *w, int K, double x, double ph) {
    for (int k = 0; k < K; k++) w[k] = w[k] + (ph == (double)(k + 1) ? x : copysign(0.0, x));
}

/* ---- phase_ratios_vertex! (src/PhaseRatios/vertices.jl:4-107) ---------------
 * ratios: CellArray over the vertex grid, element (node, phase k) at node + k*NN.
 * Loop order: offset_i (x) OUTERMOST ... offset_k innermost; half-cell test with >=;
 * NaNs (vertices with no particle in range) are left as they are. */
int jpo_phase_ratios_vertex(const jpo_grid *g, const double *const *coords, double *ratios, const double *phases, int K) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1], nz = N == 3 ? g->n[2] : 0;
    const int64_t NN = (int64_t)(nx + 1) * (ny + 1) * (N == 3 ? nz + 1 : 1);
    if (K > 64) return -1;
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t nd = 0; nd < NN; nd++) {
        int I[3] = {(int)(nd % (nx + 1)), (int)((nd / (nx + 1)) % (ny + 1)), (int)(nd / ((int64_t)(nx + 1) * (ny + 1)))};
        double xv[3] = {g->xv[0][I[0]], g->xv[1][I[1]], N == 3 ? g->xv[2][I[2]] : 0.0};
        double w[64];
        for (int k = 0; k < K; k++) w[k] = 0.0;
        for (int oi = -1; oi <= 0; oi++) {
            int ic = I[0] + oi;
            if (ic < 0 || ic >= nx) continue;
            for (int oj = -1; oj <= 0; oj++) {
                int jc = I[1] + oj;
                if (jc < 0 || jc >= ny) continue;
                for (int ok = (N == 3 ? -1 : 0); ok <= 0; ok++) {
                    int kc = N == 3 ? I[2] + ok : 0;
                    if (N == 3 && (kc < 0 || kc >= nz)) continue;
                    int cc[3] = {ic, jc, kc};
                    double di[3];
                    for (int d = 0; d < N; d++) di[d] = d_of(g->xv[d], g->uniform, cc[d]);
                    const int64_t c = ic + (int64_t)nx * (jc + (int64_t)ny * kc);
                    for (int s = 0; s < g->S; s++) {
                        const int64_t e = c + (int64_t)s * C;
                        double p[3];
                        int nan = 0, out = 0;
                        for (int d = 0; d < N; d++) { p[d] = coords[d][e]; nan |= isnan(p[d]); }
                        if (nan) continue;
                        for (int d = 0; d < N; d++) if (fabs(p[d] - xv[d]) >= di[d] / 2) { out = 1; break; }
                        if (out) continue;
                        acc_phase(w, K, bilinear_weight(N, xv, p, di), phases[e]);
                    }
                }
            }
        }
        double sum = w[0];
        for (int k = 1; k < K; k++) sum = sum + w[k];
        const double inv = 1.0 / sum;
        for (int k = 0; k < K; k++) ratios[nd + (int64_t)k * NN] = w[k] * inv;
    }
    return 0;
}

/* ---- phase_ratios_face! (src/PhaseRatios/midpoints.jl:3-82) -----------------
 * dim = 0/1/2 (:x/:y/:z).  ratios: CellArray over the face grid (n + e_dim), element
 * (face node, k) at node + k*NF.  Work-item = cell I: face I+e_dim from the cells I and
 * min(I+e_dim, n) (the last cell is visited twice, as in the reference), plus the low
 * boundary face when I[dim] == 1.  NaN (no particle in range) -> 0. */
int jpo_phase_ratios_face(const jpo_grid *g, const double *const *coords, double *ratios, const double *phases, int K, int dim) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1], nz = N == 3 ? g->n[2] : 1;
    int nf[3] = {nx + (dim == 0), ny + (dim == 1), nz + (N == 3 && dim == 2)};
    const int64_t NF = (int64_t)nf[0] * nf[1] * nf[2];
    if (K > 64 || dim < 0 || dim >= N) return -1;
    int off[3] = {dim == 0, dim == 1, dim == 2};
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c0 = 0; c0 < C; c0++) {
        int I[3] = {(int)(c0 % nx), (int)((c0 / nx) % ny), (int)(c0 / ((int64_t)nx * ny))};
        double di[3], cen[3], face[3], w[64];
        for (int d = 0; d < N; d++) {
            di[d] = d_of(g->xv[d], g->uniform, I[d]);
            cen[d] = g->xc[d][I[d]];
            face[d] = cen[d] + di[d] * (double)off[d] / 2;
        }
        for (int k = 0; k < K; k++) w[k] = 0.0;
        for (int pass = 0; pass < 2; pass++) {
            int cc[3];
            for (int d = 0; d < 3; d++) { cc[d] = I[d] + (pass ? off[d] : 0); if (cc[d] > g->n[d] - 1 && d < N) cc[d] = g->n[d] - 1; }
            if (N == 2) cc[2] = 0;
            for (int d = 0; d < N; d++) di[d] = d_of(g->xv[d], g->uniform, cc[d]);       /* `di` is reassigned (outer variable) */
            const int64_t c = cc[0] + (int64_t)nx * (cc[1] + (int64_t)ny * cc[2]);
            for (int s = 0; s < g->S; s++) {
                const int64_t e = c + (int64_t)s * C;
                double p[3];
                int nan = 0, in = 1;
                for (int d = 0; d < N; d++) { p[d] = coords[d][e]; nan |= isnan(p[d]); }
                if (nan) continue;
                for (int d = 0; d < N; d++) in &= fabs(p[d] - face[d]) <= di[d] / 2;
                if (!in) continue;
                acc_phase(w, K, bilinear_weight(N, face, p, di), phases[e]);
            }
        }
        double sum = w[0];
        for (int k = 1; k < K; k++) sum = sum + w[k];
        double inv = 1.0 / sum;
        const int64_t fo = (I[0] + off[0]) + (int64_t)nf[0] * ((I[1] + off[1]) + (int64_t)nf[1] * (I[2] + off[2]));
        for (int k = 0; k < K; k++) { double v = w[k] * inv; ratios[fo + (int64_t)k * NF] = isnan(v) ? 0.0 : v; }
        if (I[dim] == 0) {                                     /* isboundary(offsets, I): low boundary face */
            for (int d = 0; d < N; d++) face[d] = cen[d] - di[d] * (double)off[d] / 2;    /* di = last assigned above */
            for (int k = 0; k < K; k++) w[k] = 0.0;
            for (int s = 0; s < g->S; s++) {
                const int64_t e = c0 + (int64_t)s * C;
                double p[3];
                int nan = 0, in = 1;
                for (int d = 0; d < N; d++) { p[d] = coords[d][e]; nan |= isnan(p[d]); }
                if (nan) continue;
                for (int d = 0; d < N; d++) in &= fabs(p[d] - face[d]) <= di[d] / 2;
                if (!in) continue;
                acc_phase(w, K, bilinear_weight(N, face, p, di), phases[e]);
            }
            sum = w[0];
            for (int k = 1; k < K; k++) sum = sum + w[k];
            inv = 1.0 / sum;
            const int64_t fb = I[0] + (int64_t)nf[0] * (I[1] + (int64_t)nf[1] * I[2]);
            for (int k = 0; k < K; k++) { double v = w[k] * inv; ratios[fb + (int64_t)k * NF] = isnan(v) ? 0.0 : v; }
        }
    }
    return 0;
}

/* ---- phase_ratios_midpoint! (src/PhaseRatios/midpoints.jl:115-242), 3-D only --------
 * plane = 0/1/2 for :xy/:yz/:xz, offsets (1,1,0)/(0,1,1)/(1,0,1).  ratios: CellArray over the
 * midpoint grid nm = n + offsets, element (node, k) at node + k*NM.  Work-item = cell I:
 * general case accumulates the cells I + offsets.*mask, mask in ((1,0,0),(0,1,0),(0,0,1),(1,1,1)),
 * clamped to the grid (so clamped cells are visited more than once), `di` reassigned per
 * visited cell; NaN -> 0.  Boundary branch (any offsets[i]*I[i] == 1, 1-based): literal,
 * including its quirks -- the `x === false` skip never fires for Ints, and the midpoint is
 * centre - (di*offsets*flip)/2 with flip = -lastboundary_offset (0 or -1), i.e. the cell
 * CENTRE in every direction that is not the last boundary. */
static void midpoint_accumulate(const jpo_grid *g, const double *const *coords, const double *phases, int K,
                                const int *I, const int *off, const double *mid, double *w) {
    static const int MASK[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 1}};
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    for (int k = 0; k < K; k++) w[k] = 0.0;
    for (int m = 0; m < 4; m++) {
        int cc[3];
        double di[3];
        for (int d = 0; d < 3; d++) {
            cc[d] = I[d] + off[d] * MASK[m][d];
            if (cc[d] > g->n[d] - 1) cc[d] = g->n[d] - 1;
            di[d] = d_of(g->xv[d], g->uniform, cc[d]);
        }
        const int64_t c = cc[0] + (int64_t)nx * (cc[1] + (int64_t)ny * cc[2]);
        for (int s = 0; s < g->S; s++) {
            const int64_t e = c + (int64_t)s * C;
            double p[3];
            int nan = 0, in = 1;
            for (int d = 0; d < 3; d++) { p[d] = coords[d][e]; nan |= isnan(p[d]); }
            if (nan) continue;
            for (int d = 0; d < 3; d++) in &= fabs(p[d] - mid[d]) <= di[d] / 2;
            if (!in) continue;
            acc_phase(w, K, bilinear_weight(3, mid, p, di), phases[e]);
        }
    }
}

int jpo_phase_ratios_midpoint(const jpo_grid *g, const double *const *coords, double *ratios, const double *phases, int K, int plane) {
    if (g->ndim != 3 || K > 64 || plane < 0 || plane > 2) return -1;
    static const int OFF[3][3] = {{1, 1, 0}, {0, 1, 1}, {1, 0, 1}};
    const int *off = OFF[plane];
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1];
    const int nm[3] = {g->n[0] + off[0], g->n[1] + off[1], g->n[2] + off[2]};
    const int64_t NM = (int64_t)nm[0] * nm[1] * nm[2];
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t c0 = 0; c0 < C; c0++) {
        int I[3] = {(int)(c0 % nx), (int)((c0 / nx) % ny), (int)(c0 / ((int64_t)nx * ny))};
        double di[3], cen[3], mid[3], w[64];
        for (int d = 0; d < 3; d++) {
            di[d] = d_of(g->xv[d], g->uniform, I[d]);
            cen[d] = g->xc[d][I[d]];
            mid[d] = cen[d] + di[d] * (double)off[d] / 2;
        }
        midpoint_accumulate(g, coords, phases, K, I, off, mid, w);
        double sum = w[0];
        for (int k = 1; k < K; k++) sum = sum + w[k];
        double inv = 1.0 / sum;
        const int64_t mo = (I[0] + off[0]) + (int64_t)nm[0] * ((I[1] + off[1]) + (int64_t)nm[1] * (I[2] + off[2]));
        for (int k = 0; k < K; k++) { double v = w[k] * inv; ratios[mo + (int64_t)k * NM] = isnan(v) ? 0.0 : v; }
        int boundary = 0;
        for (int d = 0; d < 3; d++) boundary |= (off[d] * (I[d] + 1) == 1);
        if (boundary) {
            int ob[3];
            for (int d = 0; d < 3; d++) ob[d] = (g->n[d] == off[d] * (I[d] + 1));
            for (int d = 0; d < 3; d++) {
                const double dI = d_of(g->xv[d], g->uniform, I[d]);
                mid[d] = cen[d] - ((dI * (double)off[d]) * (double)(-ob[d])) / 2;
            }
            midpoint_accumulate(g, coords, phases, K, I, off, mid, w);
            sum = w[0];
            for (int k = 1; k < K; k++) sum = sum + w[k];
            inv = 1.0 / sum;
            for (int pass = 0; pass < 2; pass++) {
                const int64_t mb = (I[0] + pass * ob[0]) + (int64_t)nm[0] * ((I[1] + pass * ob[1]) + (int64_t)nm[1] * (I[2] + pass * ob[2]));
                for (int k = 0; k < K; k++) { double v = w[k] * inv; ratios[mb + (int64_t)k * NM] = isnan(v) ? 0.0 : v; }
            }
        }
    }
    return 0;
}

/* ---- update_cell_halo! semantics for ONE array on ONE axis (test helper) ---
 * ImplicitGlobalGrid.update_halo! with overlap 2 / halowidth 1
 * (src/CellArrays/ImplicitGlobalGrid.jl:36-41): my plane 2 -> left nbr's plane
 * n; my plane n-1 -> right nbr's plane 1 (1-based).  Pack/unpack a cell-plane
 * (all slots) of a [C*S] array of `esz`-byte elements.                       */
int jpo_plane_copy(const jpo_grid *g, void *array, void *buf, int esz, int dim, int plane0, int pack) {
    const int N = g->ndim;
    const int64_t C = NCELLS(g);
    const int nx = g->n[0], ny = g->n[1], nz = N == 3 ? g->n[2] : 1;
    int ext[3] = {nx, ny, nz};
    int64_t m = 0;
    for (int s = 0; s < g->S; s++)
        for (int k = 0; k < (dim == 2 ? 1 : ext[2]); k++)
            for (int j = 0; j < (dim == 1 ? 1 : ext[1]); j++)
                for (int i = 0; i < (dim == 0 ? 1 : ext[0]); i++) {
                    int ci[3] = {dim == 0 ? plane0 : i, dim == 1 ? plane0 : j, dim == 2 ? plane0 : k};
                    const int64_t e = ci[0] + (int64_t)nx * (ci[1] + (int64_t)ny * ci[2]) + (int64_t)s * C;
                    if (pack) memcpy((char *)buf + m * esz, (char *)array + e * esz, (size_t)esz);
                    else      memcpy((char *)array + e * esz, (char *)buf + m * esz, (size_t)esz);
                    m++;
                }
    return 0;
}
