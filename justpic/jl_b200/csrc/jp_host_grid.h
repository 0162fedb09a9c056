// jp_host_grid.h -- host-side analysis of a jp_grid_desc: classifies the
// staggered velocity grid vectors (V / G / other), checks the preconditions of
// the fast advection path, precomputes inverse spacings and packs every vector
// into one buffer.  Shared by the CUDA library (jp_ctx_create) and the CPU
// emulation used by the no-GPU tests.
#pragma once
#include <string.h>
#include <vector>
#include "../../../include/justpic_c.h"
#include "jp_core.h"

struct JpGridOffsets {
    size_t xv[3], xc[3], xvel[3][3], xg[3], ixv[3], ixg[3];
    bool has_xg[3];
};

static inline bool jp_same_vec(const double *a, int na, const double *b, int nb) {
    return a && b && na == nb && memcmp(a, b, sizeof(double) * na) == 0;
}

// returns nullptr on success, else an error message
static inline const char *jp_grid_build(const jp_grid_desc *d, JpGrid &g, std::vector<double> &h, JpGridOffsets &o) {
    if (!d) return "null grid description";
    if (d->ndim != 2 && d->ndim != 3) return "ndim must be 2 or 3";
    if (d->S < 1 || d->S > JP_MAX_SLOTS_WIDE) return "need 1 <= max_xcell <= 1024";
    const int N = d->ndim;
    for (int a = 0; a < N; a++) {
        if (d->n[a] < 2) return "need >= 2 cells per dimension";
        if (!d->xv[a] || !d->xc[a]) return "null grid vector";
        for (int b = 0; b < N; b++)
            if (!d->xvel[a][b] || d->nvel[a][b] < 2) return "bad velocity grid vector";
    }
    memset(&g, 0, sizeof(g));
    memset(&o, 0, sizeof(o));
    g.ndim = N; g.S = d->S; g.uniform = d->uniform ? 1 : 0;
    g.n[0] = d->n[0]; g.n[1] = d->n[1]; g.n[2] = N == 3 ? d->n[2] : 1;
    g.C = (int64_t)g.n[0] * g.n[1] * g.n[2];
    const double *hxg[3] = {nullptr, nullptr, nullptr};
    int fast = 1;
    for (int dim = 0; dim < N; dim++) {
        for (int c = 0; c < N; c++) {
            g.nvel[c][dim] = d->nvel[c][dim];
            if (jp_same_vec(d->xvel[c][dim], d->nvel[c][dim], d->xv[dim], d->n[dim] + 1)) g.vkind[c][dim] = 1;
            else if (d->nvel[c][dim] == d->n[dim] + 2) {
                if (!hxg[dim]) { hxg[dim] = d->xvel[c][dim]; g.vkind[c][dim] = 2; }
                else g.vkind[c][dim] = jp_same_vec(d->xvel[c][dim], d->nvel[c][dim], hxg[dim], d->n[dim] + 2) ? 2 : 0;
            } else g.vkind[c][dim] = 0;
            if (g.vkind[c][dim] == 0) fast = 0;
        }
        if (hxg[dim])
            for (int i = 0; i <= d->n[dim]; i++)
                if (!(hxg[dim][i] < d->xv[dim][i] && d->xv[dim][i] < hxg[dim][i + 1])) fast = 0;
        for (int i = 0; i < d->n[dim]; i++)
            if (!(d->xv[dim][i] < d->xv[dim][i + 1])) fast = 0;
    }
    g.fast = fast;
    // exactly-affine grid vectors (e.g. a range over a power-of-two cell count): every stored entry must
    // equal fma(i, x[1]-x[0], x[0]) bit for bit.  affine = 1: all vertex vectors; 2: ghosted-centre vectors too
    int aff_v = fast && g.uniform, aff_g = 1;
    for (int dim = 0; dim < N && aff_v; dim++) {
        const double *xv = d->xv[dim], *xg = hxg[dim];
        const double dv = xv[1] - xv[0];
        for (int i = 0; i <= d->n[dim] && aff_v; i++) if (xv[i] != fma((double)i, dv, xv[0])) aff_v = 0;
        g.aff_v0[dim] = xv[0]; g.aff_dv[dim] = dv;
        if (!xg) { aff_g = 0; continue; }
        const double dg = xg[1] - xg[0];
        for (int i = 0; i <= d->n[dim] + 1 && aff_g; i++) if (xg[i] != fma((double)i, dg, xg[0])) aff_g = 0;
        g.aff_g0[dim] = xg[0]; g.aff_dg[dim] = dg;
    }
    g.affine = aff_v ? (aff_g ? 2 : 1) : 0;
    for (int dim = 0; dim < N; dim++) {
        g.dom_lo[dim] = d->xv[dim][0]; g.dom_hi[dim] = d->xv[dim][d->n[dim]];
        g.dxv0[dim] = d->xv[dim][1] - d->xv[dim][0];
        double m = d->xv[dim][1] - d->xv[dim][0];
        for (int i = 1; i < d->n[dim]; i++) { const double q = d->xv[dim][i + 1] - d->xv[dim][i]; m = q < m ? q : m; }
        g.inv_dmin_v[dim] = 1.0 / fabs(m);
    }
    // jp_classify_fast precondition: range grid, vertices affine within 1e-6 dx, dx far above the ulp of the coordinates
    int cf = g.uniform;
    for (int dim = 0; dim < N && cf; dim++) {
        const double *xv = d->xv[dim];
        const double dx0 = xv[1] - xv[0];
        const double amax = fmax(fabs(xv[0]), fabs(xv[d->n[dim]]));
        if (!(dx0 > 0) || !(amax * 3.6e-15 <= 1e-6 * dx0)) { cf = 0; break; }
        for (int i = 0; i <= d->n[dim]; i++)
            if (!(fabs(xv[i] - (xv[0] + i * dx0)) <= 1e-6 * dx0)) { cf = 0; break; }
    }
    g.cls_fast = cf;
    h.clear();
    auto push = [&](const double *x, int n) { size_t off = h.size(); h.insert(h.end(), x, x + n); return off; };
    for (int dim = 0; dim < N; dim++) {
        o.xv[dim] = push(d->xv[dim], d->n[dim] + 1);
        o.xc[dim] = push(d->xc[dim], d->n[dim]);
        for (int c = 0; c < N; c++) o.xvel[c][dim] = push(d->xvel[c][dim], d->nvel[c][dim]);
        std::vector<double> inv(d->n[dim]);
        for (int i = 0; i < d->n[dim]; i++) inv[i] = 1.0 / (d->xv[dim][i + 1] - d->xv[dim][i]);
        o.ixv[dim] = push(inv.data(), d->n[dim]);
        g.inv_dv[dim] = 1.0 / (d->xv[dim][1] - d->xv[dim][0]);
        o.has_xg[dim] = hxg[dim] != nullptr;
        if (hxg[dim]) {
            o.xg[dim] = push(hxg[dim], d->n[dim] + 2);
            std::vector<double> ig(d->n[dim] + 1);
            for (int i = 0; i <= d->n[dim]; i++) ig[i] = 1.0 / (hxg[dim][i + 1] - hxg[dim][i]);
            o.ixg[dim] = push(ig.data(), d->n[dim] + 1);
            g.inv_dg[dim] = 1.0 / (hxg[dim][1] - hxg[dim][0]);
            g.dxg0[dim] = hxg[dim][1] - hxg[dim][0];
        }
    }
    return nullptr;
}

// point the grid tables at `base` (device or host copy of the packed buffer)
static inline void jp_grid_rebase(JpGrid &g, const JpGridOffsets &o, const double *base) {
    for (int dim = 0; dim < g.ndim; dim++) {
        g.xv[dim] = base + o.xv[dim];
        g.xc[dim] = base + o.xc[dim];
        g.ixv[dim] = base + o.ixv[dim];
        for (int c = 0; c < g.ndim; c++) g.xvel[c][dim] = base + o.xvel[c][dim];
        g.xg[dim] = o.has_xg[dim] ? base + o.xg[dim] : nullptr;
        g.ixg[dim] = o.has_xg[dim] ? base + o.ixg[dim] : nullptr;
    }
}
