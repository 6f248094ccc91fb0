/*
 * kamr_oracle.c — CPU restatement of KitAMR.jl's per-step phase-space path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (kitamr.jl_b200/,
 * libkamr) may include, link or call this file; only tests/, the smoke check
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs
 * use it, as the checker or as the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference is pure Julia + libp4est + MPI, none of which
 * exist in this image, and its own test (test/runtests.jl) asserts nothing and
 * ships no golden vectors.  This file therefore restates
 * the reference source function by function (citations below, paths relative
 * to /root/reference) in the same operation order — sequential merge-walks,
 * divisions where the reference divides, face-loop scatter — and is compiled
 * with -ffp-contract=off so no FMA is introduced that Julia would not emit.
 * Reductions are plain left-to-right sums (Julia's `sum` is pairwise/SIMD, so
 * reductions agree to rounding only).
 *
 * Data model: the flat host arrays of include/kamr.h (the drop-in boundary).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "../include/kamr.h"
#include "kamr_oracle.h"

#define EPS_KIT 1e-12                   /* src/Abstract/Types.jl:3 */
#define EPS_MACH 2.220446049250313e-16  /* Julia eps() == 2^-52, Flux/CAIDVM.jl:137 */
#define MAXD 3
#define MAXM 5

static const double PI_ = 3.14159265358979323846;

typedef struct {
    int D, K, M;            /* DIM, NDF, DIM+2 */
    const kamr_config* cfg;
    const kamr_mesh* m;
    int n_cell;
    int64_t* vs_off;        /* [n_cell+1] */
} octx;

static inline double pow2i(int e) { return ldexp(1.0, e); }

static int octx_init(octx* o, const kamr_config* cfg, const kamr_mesh* m) {
    o->D = cfg->dim; o->K = cfg->ndf; o->M = cfg->dim + 2; o->cfg = cfg; o->m = m;
    o->n_cell = m->n_local + m->n_ghost + m->n_solidnbr;
    o->vs_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(o->n_cell + 1));
    if (!o->vs_off) return 1;
    o->vs_off[0] = 0;
    for (int c = 0; c < o->n_cell; ++c) {
        int g = m->cell_grid[c];
        o->vs_off[c + 1] = o->vs_off[c] + (m->grid_off[g + 1] - m->grid_off[g]);
    }
    return 0;
}
static void octx_free(octx* o) { free(o->vs_off); }

static inline int cell_n(const octx* o, int c) {
    int g = o->m->cell_grid[c];
    return (int)(o->m->grid_off[g + 1] - o->m->grid_off[g]);
}
static inline const int8_t* cell_level(const octx* o, int c) {
    return o->m->v_level + o->m->grid_off[o->m->cell_grid[c]];
}
static inline const double* cell_weight(const octx* o, int c) {
    return o->m->v_weight + o->m->grid_off[o->m->cell_grid[c]];
}
static inline const double* cell_vmid(const octx* o, int c) { /* plane d at +d*n */
    return o->m->v_mid + o->m->grid_off[o->m->cell_grid[c]] * o->D;
}
static inline double* cell_df(const octx* o, orc_state* s, int c) { return s->df + o->vs_off[c] * o->K; }
static inline double* cell_sdf(const octx* o, orc_state* s, int c) { return s->sdf + o->vs_off[c] * o->K * o->D; }
static inline double* cell_flux(const octx* o, orc_state* s, int c) { return s->flux + o->vs_off[c] * o->K; }

/* ------------------------------------------------------------------ kinetics */

/* lib/KitCore/2D.jl:9-16, 3D.jl:12-20 */
void orc_get_prim(int D, const double* w, double gamma, double* prim) {
    if (D == 2) {
        prim[0] = w[0];
        prim[1] = w[1] / w[0];
        prim[2] = w[2] / w[0];
        prim[3] = 0.5 * w[0] / (gamma - 1.0) / (w[3] - 0.5 * (w[1] * w[1] + w[2] * w[2]) / w[0]);
    } else {
        prim[0] = w[0];
        prim[1] = w[1] / w[0];
        prim[2] = w[2] / w[0];
        prim[3] = w[3] / w[0];
        prim[4] = 0.5 * w[0] / (gamma - 1.0) /
                  (w[4] - 0.5 * (w[1] * w[1] + w[2] * w[2] + w[3] * w[3]) / w[0]);
    }
}
/* lib/KitCore/2D.jl:1-8, 3D.jl:1-11 */
void orc_get_conserved(int D, const double* prim, double gamma, double* w) {
    if (D == 2) {
        w[0] = prim[0];
        w[1] = prim[0] * prim[1];
        w[2] = prim[0] * prim[2];
        w[3] = 0.5 * prim[0] / prim[3] / (gamma - 1.0) +
               0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2]);
    } else {
        w[0] = prim[0];
        w[1] = prim[0] * prim[1];
        w[2] = prim[0] * prim[2];
        w[3] = prim[0] * prim[3];
        w[4] = 0.5 * prim[0] / prim[4] / (gamma - 1.0) +
               0.5 * prim[0] * (prim[1] * prim[1] + prim[2] * prim[2] + prim[3] * prim[3]);
    }
}
/* src/Gas/Model.jl:14 */
double orc_get_tau(int D, const double* prim, double mu, double omega) {
    return mu * 2.0 * pow(prim[D + 1], 1 - omega) / prim[0];
}
/* lib/KitCore/2D2F.jl:14-21 (scalar form), 3D1F.jl:15-23 */
static inline void maxwell_point(int D, int K, const double* v, const double* prim, double Kin, double* out) {
    if (D == 2) {
        double du = v[0] - prim[1], dv = v[1] - prim[2];
        double h = prim[0] * (prim[3] / PI_) * exp(-prim[3] * (du * du + dv * dv));
        out[0] = h;
        if (K > 1) out[1] = h * Kin / (2.0 * prim[3]);
    } else {
        double du = v[0] - prim[1], dv = v[1] - prim[2], dw = v[2] - prim[3];
        out[0] = prim[0] * pow(prim[4] / PI_, 3.0 / 2.0) * exp(-prim[4] * (du * du + dv * dv + dw * dw));
    }
}
/* lib/KitCore/2D2F.jl:45-67, 3D1F.jl:24-40: F+ for one point given H,B (or M) */
static inline void shakhov_point(int D, int K, const double* v, const double* F, const double* prim,
                                 const double* qf, double Pr, double Kin, double* out) {
    if (D == 2) {
        double du = v[0] - prim[1], dv = v[1] - prim[2];
        double lam = prim[3];
        double c0 = 0.8 * (1 - Pr) * lam * lam / prim[0] * (du * qf[0] + dv * qf[1]);
        out[0] = c0 * (2 * lam * (du * du + dv * dv) + Kin - 5) * F[0];
        if (K > 1) out[1] = c0 * (2 * lam * (du * du + dv * dv) + Kin - 3) * F[1];
    } else {
        double du = v[0] - prim[1], dv = v[1] - prim[2], dw = v[2] - prim[3];
        double lam = prim[4];
        out[0] = 0.8 * (1 - Pr) * lam * lam / prim[0] * (du * qf[0] + dv * qf[1] + dw * qf[2]) *
                 (2 * lam * (du * du + dv * dv + dw * dw) - 5) * F[0];
    }
}
/* micro_to_macro: lib/KitCore/2D2F.jl:119-126, 3D1F.jl:109-124.
 * df/vmid given as planes with stride n and an optional index list (masked views). */
static void micro_to_macro_idx(int D, int K, int cnt, const int* idx, const double* micro /*cnt x K, col-major*/,
                               int n, const double* vmid, const double* weight, double* w) {
    double s[MAXM] = {0, 0, 0, 0, 0};
    double sb = 0.0;
    for (int a = 0; a < cnt; ++a) {
        int i = idx ? idx[a] : a;
        double wt = weight[i];
        double h = micro[a];
        if (D == 2) {
            double u = vmid[i], v = vmid[n + i];
            s[0] += wt * h;
            s[1] += wt * u * h;
            s[2] += wt * v * h;
            double b = (K > 1) ? micro[cnt + a] : 0.0;
            sb += wt * ((u * u + v * v) * h + b);
        } else {
            double u = vmid[i], v = vmid[n + i], ww = vmid[2 * n + i];
            s[0] += wt * h;
            s[1] += wt * u * h;
            s[2] += wt * v * h;
            s[3] += wt * ww * h;
            sb += wt * (u * u + v * v + ww * ww) * h;
        }
    }
    for (int d = 0; d <= D; ++d) w[d] = s[d];
    w[D + 1] = 0.5 * sb;
}
/* heat flux: lib/KitCore/2D2F.jl:68-88, 3D1F.jl:41-72 */
static void heat_flux(int D, int K, int n, const double* vmid, const double* df, const double* prim,
                      const double* weight, double* q) {
    if (D == 2) {
        double a1 = 0, b1 = 0, a2 = 0, b2 = 0;
        for (int i = 0; i < n; ++i) {
            double du = vmid[i] - prim[1], dv = vmid[n + i] - prim[2];
            double c2 = du * du + dv * dv;
            a1 += weight[i] * du * c2 * df[i];
            a2 += weight[i] * dv * c2 * df[i];
            if (K > 1) {
                b1 += weight[i] * du * df[n + i];
                b2 += weight[i] * dv * df[n + i];
            }
        }
        q[0] = 0.5 * (a1 + b1);
        q[1] = 0.5 * (a2 + b2);
    } else {
        double a1 = 0, a2 = 0, a3 = 0;
        for (int i = 0; i < n; ++i) {
            double du = vmid[i] - prim[1], dv = vmid[n + i] - prim[2], dw = vmid[2 * n + i] - prim[3];
            double c2 = du * du + dv * dv + dw * dw;
            a1 += weight[i] * du * c2 * df[i];
            a2 += weight[i] * dv * c2 * df[i];
            a3 += weight[i] * dw * c2 * df[i];
        }
        q[0] = 0.5 * a1; q[1] = 0.5 * a2; q[2] = 0.5 * a3;
    }
}

/* exported point-wise helpers for unit tests */
void orc_discrete_maxwell(int D, int K, int n, const double* vmid, const double* prim, double Kin, double* F) {
    for (int i = 0; i < n; ++i) {
        double v[MAXD], out[2];
        for (int d = 0; d < D; ++d) v[d] = vmid[d * n + i];
        maxwell_point(D, K, v, prim, Kin, out);
        for (int k = 0; k < K; ++k) F[k * n + i] = out[k];
    }
}
void orc_shakhov_part(int D, int K, int n, const double* vmid, const double* F, const double* prim,
                      const double* qf, double Pr, double Kin, double* Fp) {
    for (int i = 0; i < n; ++i) {
        double v[MAXD], f[2], out[2];
        for (int d = 0; d < D; ++d) v[d] = vmid[d * n + i];
        for (int k = 0; k < K; ++k) f[k] = F[k * n + i];
        shakhov_point(D, K, v, f, prim, qf, Pr, Kin, out);
        for (int k = 0; k < K; ++k) Fp[k * n + i] = out[k];
    }
}
void orc_micro_to_macro(int D, int K, int n, const double* vmid, const double* df, const double* weight, double* w) {
    micro_to_macro_idx(D, K, n, NULL, df, n, vmid, weight, w);
}
void orc_heat_flux(int D, int K, int n, const double* vmid, const double* df, const double* prim,
                   const double* weight, double* q) {
    heat_flux(D, K, n, vmid, df, prim, weight, q);
}

/* ------------------------------------------------------------------ pair map */
/* Coverage of grid a's points by grid b's points as visited by the reference's
 * merge-walks (Flux/Slope.jl:29-64, Flux/Flux.jl:155-279): start[i] = first b
 * index matched with a's point i; start[n_a] = n_b.  Returns 0 ok, 1 if the
 * walk runs off either grid (grids do not tile the same domain). */
int orc_pair_map(int D, int n_a, const int8_t* lev_a, int n_b, const int8_t* lev_b, int32_t* start) {
    int index = 0;
    double flag = 0.0;
    for (int i = 0; i < n_a; ++i) {
        if (index >= n_b) return 1;
        start[i] = index;
        if (lev_a[i] == lev_b[index]) {
            index += 1;
        } else if (lev_a[i] < lev_b[index]) {
            while (flag != 1.0) {
                if (index >= n_b) return 1;
                flag += 1 / pow2i(D * (lev_b[index] - lev_a[i]));
                index += 1;
            }
            flag = 0.0;
        } else {
            flag += 1 / pow2i(D * (lev_a[i] - lev_b[index]));
            if (flag == 1.0) { index += 1; flag = 0.0; }
        }
    }
    start[n_a] = n_b;
    return index == n_b ? 0 : 1;
}

/* ------------------------------------------------------------------ slopes */

/* minmod, Flux/Slope.jl:20-24 */
static inline double sgn(double x) { return (x > 0) - (x < 0); }
static inline double minmod(double sL, double sR) {
    double SL = fabs(sL), SR = fabs(sR);
    return 0.5 * (sgn(sL) + sgn(sR)) * fmin(SL, SR);
}

/* diff_vs!, Flux/Slope.jl:29-64 ; diff_vs_transverse!, :278-333 when dm != NULL */
static void diff_vs(const octx* o, orc_state* st, int c, int cn, double dsL, const double* dm, double* sL) {
    const int D = o->D, K = o->K;
    const int n = cell_n(o, c), nn = cell_n(o, cn);
    const int8_t* level = cell_level(o, c);
    const int8_t* level_n = cell_level(o, cn);
    const double* df = cell_df(o, st, c);
    const double* dfn = cell_df(o, st, cn);
    const double* sdfn = cell_sdf(o, st, cn);
    int index = 0;
    double flag = 0.0;
    for (int i = 0; i < n; ++i) {
        if (level[i] == level_n[index]) {
            for (int j = 0; j < K; ++j) {
                double proj = dfn[j * nn + index];
                if (dm) for (int t = 0; t < D; ++t) proj += dm[t] * sdfn[(t * K + j) * nn + index];
                sL[j * n + i] += (df[j * n + i] - proj) / dsL;
            }
            index += 1;
        } else if (level[i] < level_n[index]) {
            while (flag != 1.0) {
                for (int j = 0; j < K; ++j) {
                    double proj = dfn[j * nn + index];
                    if (dm) for (int t = 0; t < D; ++t) proj += dm[t] * sdfn[(t * K + j) * nn + index];
                    sL[j * n + i] += (df[j * n + i] - proj) / pow2i(D * (level_n[index] - level[i])) / dsL;
                }
                flag += 1 / pow2i(D * (level_n[index] - level[i]));
                index += 1;
            }
            flag = 0.0;
        } else {
            for (int j = 0; j < K; ++j) {
                double proj = dfn[j * nn + index];
                if (dm) for (int t = 0; t < D; ++t) proj += dm[t] * sdfn[(t * K + j) * nn + index];
                sL[j * n + i] += (df[j * n + i] - proj) / dsL;
            }
            flag += 1 / pow2i(D * (level[i] - level_n[index]));
            if (flag == 1.0) { index += 1; flag = 0.0; }
        }
    }
}

typedef struct { int cnt; const int32_t* ids; } nlist;
static inline nlist nb_list(const octx* o, int c, int face) {
    const kamr_mesh* m = o->m;
    int e = c * 2 * o->D + face;
    nlist l = { m->nb_off[e + 1] - m->nb_off[e], m->nb_ids + m->nb_off[e] };
    return l;
}

/* _ps_has_transverse_offset, Flux/Slope.jl:177-201 */
static int has_transverse_offset(const octx* o, int c, nlist nb, int dir) {
    const int D = o->D;
    if (D == 1 || nb.cnt == 0) return 0;
    for (int t = 0; t < D; ++t) {
        if (t == dir) continue;
        double avg = 0.0;
        for (int j = 0; j < nb.cnt; ++j) avg += o->m->mid[nb.ids[j] * D + t];
        avg /= nb.cnt;
        if (avg != o->m->mid[c * D + t]) return 1;
    }
    return 0;
}

/* update_slope_bound_vs!, Slope.jl:68-86 (transverse: :428-452, always projects) */
static void slope_bound_vs(const octx* o, orc_state* st, int c, nlist nb, double ds, int dir, int transverse,
                           double* ws) {
    const int D = o->D, K = o->K, n = cell_n(o, c);
    memset(ws, 0, sizeof(double) * (size_t)n * K);
    for (int j = 0; j < nb.cnt; ++j) {
        if (transverse) {
            double dm[MAXD];
            for (int t = 0; t < D; ++t)
                dm[t] = (t == dir) ? 0.0 : (o->m->mid[c * D + t] - o->m->mid[nb.ids[j] * D + t]);
            diff_vs(o, st, c, nb.ids[j], ds, dm, ws);
        } else {
            diff_vs(o, st, c, nb.ids[j], ds, NULL, ws);
        }
    }
    double* sdf = cell_sdf(o, st, c);
    for (int k = 0; k < K; ++k)
        for (int i = 0; i < n; ++i) sdf[(dir * K + k) * n + i] = ws[k * n + i] / nb.cnt;
}

/* update_slope_inner_vs!, Slope.jl:90-116 (transverse: :339-383, per-side projection) */
static void slope_inner_vs(const octx* o, orc_state* st, int c, nlist L, nlist R, double dsL, double dsR, int dir,
                           int transverse, double* wsL, double* wsR) {
    const int D = o->D, K = o->K, n = cell_n(o, c);
    memset(wsL, 0, sizeof(double) * (size_t)n * K);
    memset(wsR, 0, sizeof(double) * (size_t)n * K);
    int projL = transverse ? has_transverse_offset(o, c, L, dir) : 0;
    int projR = transverse ? has_transverse_offset(o, c, R, dir) : 0;
    double dm[MAXD];
    for (int j = 0; j < L.cnt; ++j) {
        if (projL) {
            for (int t = 0; t < D; ++t)
                dm[t] = (t == dir) ? 0.0 : (o->m->mid[c * D + t] - o->m->mid[L.ids[j] * D + t]);
            diff_vs(o, st, c, L.ids[j], dsL, dm, wsL);
        } else diff_vs(o, st, c, L.ids[j], dsL, NULL, wsL);
    }
    for (int j = 0; j < R.cnt; ++j) {
        if (projR) {
            for (int t = 0; t < D; ++t)
                dm[t] = (t == dir) ? 0.0 : (o->m->mid[c * D + t] - o->m->mid[R.ids[j] * D + t]);
            diff_vs(o, st, c, R.ids[j], dsR, dm, wsR);
        } else diff_vs(o, st, c, R.ids[j], dsR, NULL, wsR);
    }
    double* sdf = cell_sdf(o, st, c);
    for (int k = 0; k < K; ++k)
        for (int i = 0; i < n; ++i)
            sdf[(dir * K + k) * n + i] = minmod(wsL[k * n + i] / L.cnt, wsR[k * n + i] / R.cnt);
}

/* the 15 update_slope! methods, Slope.jl:458-771 */
static int update_slope(const octx* o, orc_state* st, int c, int dir, double* wsL, double* wsR) {
    const kamr_mesh* m = o->m;
    const int D = o->D, K = o->K, n = cell_n(o, c);
    int sL = m->nb_state[c * 2 * D + 2 * dir], sR = m->nb_state[c * 2 * D + 2 * dir + 1];
    nlist L = nb_list(o, c, 2 * dir), R = nb_list(o, c, 2 * dir + 1);
    const double* mid = m->mid;
    if (sL == 1 && sR == 1) { /* :458-488 */
        int solidL = m->bound_enc[L.ids[0]] < 0, solidR = m->bound_enc[R.ids[0]] < 0;
        if (solidL && solidR) {
            double* sdf = cell_sdf(o, st, c);
            for (int k = 0; k < K; ++k)
                for (int i = 0; i < n; ++i) sdf[(dir * K + k) * n + i] = 0.0;
        } else if (solidL) {
            double ds = mid[c * D + dir] - mid[R.ids[0] * D + dir];
            slope_bound_vs(o, st, c, R, ds, dir, 0, wsL);
        } else if (solidR) {
            double ds = mid[c * D + dir] - mid[L.ids[0] * D + dir];
            slope_bound_vs(o, st, c, L, ds, dir, 0, wsL);
        } else {
            double ds = m->ds[c * D + dir];
            slope_inner_vs(o, st, c, L, R, ds, -ds, dir, 0, wsL, wsR);
        }
        return 0;
    }
    if (sL == 0 && sR == 0) return 1; /* no such method in the reference */
    if (sL == 0) { /* :653-668, :716-731, :735-750 */
        double ds = mid[c * D + dir] - mid[R.ids[0] * D + dir];
        slope_bound_vs(o, st, c, R, ds, dir, 0, wsL);
        return 0;
    }
    if (sR == 0) { /* :672-687, :691-712, :754-771 */
        double ds = mid[c * D + dir] - mid[L.ids[0] * D + dir];
        slope_bound_vs(o, st, c, L, ds, dir, 0, wsL);
        return 0;
    }
    /* inner, :492-649: multiplier 1 (same), 0.75 (finer list), 1.5 (coarser) */
    double ds = m->ds[c * D + dir];
    double fL = (sL == 1) ? 1.0 : (sL == -1 ? 1.5 : 0.75);
    double fR = (sR == 1) ? 1.0 : (sR == -1 ? 1.5 : 0.75);
    double dsL = (sL == 1) ? ds : fL * ds;
    double dsR = (sR == 1) ? -ds : -fR * ds;
    slope_inner_vs(o, st, c, L, R, dsL, dsR, dir, 0, wsL, wsR);
    return 0;
}

static inline int skip_cell(const octx* o, int c) { return o->m->bound_enc[c] < 0; }

/* update_slope_level!, Slope.jl:977-1019 */
static int update_slope_level(const octx* o, orc_state* st, int Lv, double* wsL, double* wsR) {
    for (int c = 0; c < o->m->n_local; ++c) {
        if (skip_cell(o, c) || o->m->ps_level[c] != Lv) continue;
        for (int dir = 0; dir < o->D; ++dir)
            if (update_slope(o, st, c, dir, wsL, wsR)) return 1;
    }
    return 0;
}

/* update_slope_transverse_level!, Slope.jl:849-945 */
static int update_slope_transverse_level(const octx* o, orc_state* st, int Lv, double* wsL, double* wsR) {
    const kamr_mesh* m = o->m;
    const int D = o->D;
    for (int c = 0; c < m->n_local; ++c) {
        if (skip_cell(o, c) || m->ps_level[c] != Lv) continue;
        for (int dir = 0; dir < D; ++dir) { /* pass (1) */
            int sL = m->nb_state[c * 2 * D + 2 * dir], sR = m->nb_state[c * 2 * D + 2 * dir + 1];
            if (sL == -1 || sR == -1) continue;
            if (update_slope(o, st, c, dir, wsL, wsR)) return 1;
        }
        if (D == 1) continue;
        for (int dir = 0; dir < D; ++dir) { /* pass (2) */
            int sL = m->nb_state[c * 2 * D + 2 * dir], sR = m->nb_state[c * 2 * D + 2 * dir + 1];
            nlist L = nb_list(o, c, 2 * dir), R = nb_list(o, c, 2 * dir + 1);
            if (sL != 0 && sR != 0) {
                if (!(sL == -1 || sR == -1)) continue;
                double dsL = m->mid[c * D + dir] - m->mid[L.ids[0] * D + dir];
                double dsR = m->mid[c * D + dir] - m->mid[R.ids[0] * D + dir];
                slope_inner_vs(o, st, c, L, R, dsL, dsR, dir, 1, wsL, wsR);
                if (m->bound_enc[L.ids[0]] < 0) slope_bound_vs(o, st, c, R, dsR, dir, 1, wsL);
                else if (m->bound_enc[R.ids[0]] < 0) slope_bound_vs(o, st, c, L, dsL, dir, 1, wsL);
            } else if (sR == 0 && sL == -1) {
                if (m->bound_enc[L.ids[0]] < 0) continue;
                double dsL = m->mid[c * D + dir] - m->mid[L.ids[0] * D + dir];
                slope_bound_vs(o, st, c, L, dsL, dir, 1, wsL);
            } else if (sL == 0 && sR == -1) {
                if (m->bound_enc[R.ids[0]] < 0) continue;
                double dsR = m->mid[c * D + dir] - m->mid[R.ids[0] * D + dir];
                slope_bound_vs(o, st, c, R, dsR, dir, 1, wsL);
            }
        }
    }
    return 0;
}

/* update_macro_slope!, Slope.jl:1022-1036 */
static void update_macro_slope(const octx* o, orc_state* st) {
    const int D = o->D, K = o->K, M = o->M;
    for (int c = 0; c < o->m->n_local; ++c) {
        if (skip_cell(o, c)) continue;
        int n = cell_n(o, c);
        for (int dir = 0; dir < D; ++dir)
            micro_to_macro_idx(D, K, n, NULL, cell_sdf(o, st, c) + (size_t)dir * K * n, n, cell_vmid(o, c),
                               cell_weight(o, c), st->sw + (size_t)c * M * D + dir * M);
    }
}

/* one level of slope! — exported so multi-rank tests can interleave the
 * per-level halo exchange exactly as Slope.jl:1055-1066 does */
int orc_slope_level(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, int Lv, int transverse) {
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    int maxn = 1;
    for (int c = 0; c < m->n_local; ++c) if (cell_n(&o, c) > maxn) maxn = cell_n(&o, c);
    double* wsL = (double*)malloc(sizeof(double) * (size_t)maxn * o.K);
    double* wsR = (double*)malloc(sizeof(double) * (size_t)maxn * o.K);
    int rc = transverse ? update_slope_transverse_level(&o, st, Lv, wsL, wsR)
                        : update_slope_level(&o, st, Lv, wsL, wsR);
    free(wsL); free(wsR);
    octx_free(&o);
    return rc;
}
int orc_macro_slope(const kamr_config* cfg, const kamr_mesh* m, orc_state* st) {
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    update_macro_slope(&o, st);
    octx_free(&o);
    return 0;
}
/* slope!, Slope.jl:1047-1070 (single address space: the exchanges are no-ops) */
int orc_slope(const kamr_config* cfg, const kamr_mesh* m, orc_state* st) {
    int rc = orc_slope_level(cfg, m, st, m->ps_minlevel, 0);
    if (rc) return rc;
    for (int L = m->ps_minlevel + 1; L <= m->ps_maxlevel; ++L) {
        rc = orc_slope_level(cfg, m, st, L, 1);
        if (rc) return rc;
    }
    return orc_macro_slope(cfg, m, st);
}

/* ------------------------------------------------------------------ flux */

/* face_area, Flux/Flux.jl:10-15 */
static double face_area(const octx* o, int c, int dir) {
    const double* ds = o->m->ds + (size_t)c * o->D;
    if (o->D == 2) return ds[dir == 0 ? 1 : 0];
    if (dir == 0) return ds[1] * ds[2];
    if (dir == 1) return ds[0] * ds[2];
    return ds[0] * ds[1];
}

typedef struct {
    int cnt; int* idx;       /* masked point indices (ascending) */
    double* micro;           /* cnt x K col-major */
} facevs;

/* positivity_preserving_reconstruct, Flux/CAIDVM.jl:127-141; unlimited branch :110-111 */
static void reconstruct(const octx* o, orc_state* st, int c, const double* ps_mid, const double* fmid, double dt,
                        int dir, facevs* fv, int mode /*0 limited,1 unlimited,2 none*/) {
    const int D = o->D, K = o->K, n = cell_n(o, c);
    const double* vm = cell_vmid(o, c);
    const double* df = cell_df(o, st, c);
    const double* sdf = cell_sdf(o, st, c);
    const double* ds = o->m->ds + (size_t)c * D;
    for (int a = 0; a < fv->cnt; ++a) {
        int i = fv->idx[a];
        double dx[MAXD];
        for (int j = 0; j < D; ++j) dx[j] = fmid[j] - vm[j * n + i] * dt - ps_mid[j];
        double vn = vm[dir * n + i];
        for (int k = 0; k < K; ++k) {
            double f = df[k * n + i];
            if (mode == 2) { fv->micro[k * fv->cnt + a] = f * vn; continue; }
            double s_abs = 0.0, s_dx = 0.0;
            for (int t = 0; t < D; ++t) {
                double s = sdf[(t * K + k) * n + i];
                s_abs += ds[t] * fabs(s);
                s_dx += dx[t] * s;
            }
            if (mode == 1) fv->micro[k * fv->cnt + a] = (f + s_dx) * vn;
            else fv->micro[k * fv->cnt + a] =
                     (f + fmin(fabs((f - EPS_MACH) / (0.5 * s_abs + EPS_KIT)), 1.) * s_dx) * vn;
        }
    }
}

/* make_face_vs masks, Flux/Flux.jl:349-424: here rot*v<=0 ; there rot*v>0 */
static void make_mask(const octx* o, int c, int dir, double rot, int there, facevs* fv) {
    const int n = cell_n(o, c);
    const double* vm = cell_vmid(o, c) + (size_t)dir * n;
    fv->cnt = 0;
    for (int i = 0; i < n; ++i) {
        double x = rot * vm[i];
        int sel = there ? (x > 0.) : (x <= 0.);
        if (sel) fv->idx[fv->cnt++] = i;
    }
}

/* update_micro_flux!, Flux/Flux.jl:151-344 (local there with/without write-back; ghost there) */
static void update_micro_flux(const octx* o, orc_state* st, int here, int there, const facevs* hv, const facevs* tv,
                              double area, int write_there) {
    const int D = o->D, K = o->K;
    const int n = cell_n(o, here), nn = cell_n(o, there);
    const int8_t* level = cell_level(o, here);
    const int8_t* level_n = cell_level(o, there);
    double* flux = cell_flux(o, st, here);
    double* flux_n = cell_flux(o, st, there);
    int index = 0, j = 0, index_n = 0, hp = 0; /* hp walks hv->idx to recover heavi[i] */
    double flag = 0.;
#define HM(k) (hv->micro[(k) * hv->cnt + index] * area)
#define TM(k) (tv->micro[(k) * tv->cnt + index_n] * area)
    for (int i = 0; i < n; ++i) {
        int heavi = (hp < hv->cnt && hv->idx[hp] == i);
        if (heavi) {
            hp++;
            for (int ii = 0; ii < K; ++ii) flux[ii * n + i] += HM(ii);
            if (level[i] == level_n[j]) {
                if (write_there) for (int ii = 0; ii < K; ++ii) flux_n[ii * nn + j] -= HM(ii);
                j += 1;
            } else if (level[i] < level_n[j]) {
                while (flag != 1.0) {
                    if (write_there) for (int ii = 0; ii < K; ++ii) flux_n[ii * nn + j] -= HM(ii);
                    flag += 1 / pow2i(D * (level_n[j] - level[i]));
                    j += 1;
                }
                flag = 0.0;
            } else {
                if (write_there)
                    for (int ii = 0; ii < K; ++ii) flux_n[ii * nn + j] -= HM(ii) / pow2i(D * (level[i] - level_n[j]));
                flag += 1 / pow2i(D * (level[i] - level_n[j]));
                if (flag == 1.0) { j += 1; flag = 0.0; }
            }
            index += 1;
        } else {
            if (level[i] == level_n[j]) {
                for (int ii = 0; ii < K; ++ii) {
                    if (write_there) flux_n[ii * nn + j] -= TM(ii);
                    flux[ii * n + i] += TM(ii);
                }
                j += 1; index_n += 1;
            } else if (level[i] < level_n[j]) {
                while (flag != 1.0) {
                    for (int ii = 0; ii < K; ++ii) {
                        if (write_there) flux_n[ii * nn + j] -= TM(ii);
                        flux[ii * n + i] += TM(ii) / pow2i(D * (level_n[j] - level[i]));
                    }
                    flag += 1 / pow2i(D * (level_n[j] - level[i]));
                    j += 1; index_n += 1;
                }
                flag = 0.0;
            } else {
                for (int ii = 0; ii < K; ++ii) flux[ii * n + i] += TM(ii);
                flag += 1 / pow2i(D * (level[i] - level_n[j]));
                if (flag == 1.0) {
                    if (write_there) for (int ii = 0; ii < K; ++ii) flux_n[ii * nn + j] -= TM(ii);
                    j += 1; index_n += 1; flag = 0.0;
                }
            }
        }
    }
#undef HM
#undef TM
}

/* flux!(F, face::FullFace|FluxData), Flux/Flux.jl:28-59,94-111 with calc_flux(CAIDVM) CAIDVM.jl:99-121 */
static void flux_inner_face(const octx* o, orc_state* st, int f, double dt, facevs* hv, facevs* tv) {
    const kamr_mesh* m = o->m;
    const int D = o->D, K = o->K, M = o->M;
    int here = m->face_here[f], there = m->face_there[f], dir = m->face_dir[f], kind = m->face_kind[f];
    double rot = m->face_rot[f];
    const double* fmid = m->face_mid + (size_t)f * D;
    make_mask(o, here, dir, rot, 0, hv);
    make_mask(o, there, dir, rot, 1, tv);
    int there_solid = m->bound_enc[there] < 0;
    reconstruct(o, st, here, m->mid + (size_t)here * D, fmid, dt, dir, hv, there_solid ? 1 : 0);
    /* there side of a solid / SolidNeighbor face: CAIDVM takes f_wall*v_n (CAIDVM.jl:111), DVM reconstructs with the
     * SolidNeighbor's own slopes, (there_df + ndx.there_sdf)*v_n (DVM.jl:91) */
    reconstruct(o, st, there, m->face_there_mid + (size_t)f * D, fmid, dt, dir, tv,
                there_solid ? (o->cfg->flux_type == KAMR_FLUX_DVM ? 1 : 2) : 0);
    double area = face_area(o, here, dir);
    if (kind == KAMR_FACE_HANGING) area = area / pow2i(D - 1) * rot; /* Flux.jl:84-86 */
    else area = area * rot;                                          /* Flux.jl:87-89 */
    if (o->cfg->flux_type == KAMR_FLUX_CAIDVM) {
        double fw[MAXM], fw2[MAXM];
        micro_to_macro_idx(D, K, hv->cnt, hv->idx, hv->micro, cell_n(o, here), cell_vmid(o, here),
                           cell_weight(o, here), fw);
        micro_to_macro_idx(D, K, tv->cnt, tv->idx, tv->micro, cell_n(o, there), cell_vmid(o, there),
                           cell_weight(o, there), fw2);
        /* update_macro_flux!, Flux.jl:116-136 */
        int there_local_fluid = (there < m->n_local) && m->bound_enc[there] >= 0;
        for (int q = 0; q < M; ++q) {
            double v = (fw[q] + fw2[q]) * area;
            st->mflux[(size_t)here * M + q] += v;
            if (there_local_fluid) st->mflux[(size_t)there * M + q] -= v;
        }
    }
    int write_there = (there < m->n_local) && m->bound_enc[there] >= 0;
    update_micro_flux(o, st, here, there, hv, tv, area, write_there);
}

/* calc_domain_flux(CAIDVM, ...) CAIDVM.jl:4-97 + update_domain_flux! Flux.jl:64-82 */
static void flux_domain_face(const octx* o, orc_state* st, int f, double dt, facevs* hv, facevs* tv) {
    const kamr_mesh* m = o->m;
    const int D = o->D, K = o->K, M = o->M;
    int c = m->face_here[f], dir = m->face_dir[f], b = m->face_there[f];
    double rot = m->face_rot[f];
    const double* fmid = m->face_mid + (size_t)f * D;
    const double* pmid = m->mid + (size_t)c * D;
    const int n = cell_n(o, c);
    const double* vm = cell_vmid(o, c);
    const double* wt = cell_weight(o, c);
    const double* df = cell_df(o, st, c);
    const double* sdf = cell_sdf(o, st, c);
    int bct = m->bc_type[b];
    double bc[MAXM];
    for (int q = 0; q < M; ++q) bc[q] = m->bc_prim[(size_t)b * M + q];
    make_mask(o, c, dir, rot, 0, hv);
    /* nheavi = !heavi */
    tv->cnt = 0;
    { int hp = 0; for (int i = 0; i < n; ++i) { if (hp < hv->cnt && hv->idx[hp] == i) hp++; else tv->idx[tv->cnt++] = i; } }
    if (bct == KAMR_BC_UNIFORM_OUTFLOW) { /* CAIDVM.jl:53-67 */
        reconstruct(o, st, c, pmid, fmid, dt, dir, hv, 2);
        reconstruct(o, st, c, pmid, fmid, dt, dir, tv, 2);
    } else if (bct == KAMR_BC_INTERPOLATED_OUTFLOW) { /* CAIDVM.jl:71-97 */
        double tmid[MAXD];
        for (int j = 0; j < D; ++j) tmid[j] = 2.0 * fmid[j] - pmid[j];
        reconstruct(o, st, c, pmid, fmid, dt, dir, hv, 1);
        for (int a = 0; a < tv->cnt; ++a) {
            int i = tv->idx[a];
            double vn = vm[dir * n + i];
            for (int k = 0; k < K; ++k) {
                double tdf = df[k * n + i] + (tmid[dir] - pmid[dir]) * sdf[(dir * K + k) * n + i];
                double s_dx = 0.0;
                for (int t = 0; t < D; ++t) s_dx += (fmid[t] - vm[t * n + i] * dt - tmid[t]) * sdf[(t * K + k) * n + i];
                tv->micro[k * tv->cnt + a] = (tdf + s_dx) * vn;
            }
        }
    } else { /* Maxwellian wall :4-25, SuperSonicInflow :29-49 */
        reconstruct(o, st, c, pmid, fmid, dt, dir, hv, 1);
        if (bct == KAMR_BC_MAXWELLIAN) { /* calc_ρw, Theory/Math.jl:251-282 */
            double SF = 0.0, SG = 0.0;
            for (int a = 0; a < hv->cnt; ++a) {
                int i = hv->idx[a];
                double vn = vm[dir * n + i];
                /* SF uses the reconstructed h component, Theory/Math.jl:261 */
                double f = df[i], s_dx = 0.0;
                for (int t = 0; t < D; ++t) s_dx += (fmid[t] - vm[t * n + i] * dt - pmid[t]) * sdf[(t * K + 0) * n + i];
                SF += wt[i] * vn * (f + s_dx);
            }
            for (int a = 0; a < tv->cnt; ++a) {
                int i = tv->idx[a];
                double c2 = 0.0;
                for (int t = 0; t < D; ++t) { double dd = vm[t * n + i] - bc[1 + t]; c2 += dd * dd; }
                SG += wt[i] * vm[dir * n + i] * exp(-bc[D + 1] * c2);
            }
            if (D == 2) SG = bc[3] / PI_ * SG; else SG = pow(bc[4] / PI_, 3.0 / 2.0) * SG;
            bc[0] = -SF / SG;
        }
        for (int a = 0; a < tv->cnt; ++a) {
            int i = tv->idx[a];
            double v[MAXD], out[2];
            for (int t = 0; t < D; ++t) v[t] = vm[t * n + i];
            maxwell_point(D, K, v, bc, o->cfg->K, out);
            for (int k = 0; k < K; ++k) tv->micro[k * tv->cnt + a] = out[k] * vm[dir * n + i];
        }
    }
    double area = rot * face_area(o, c, dir);
    if (o->cfg->flux_type == KAMR_FLUX_CAIDVM) {
        double fw[MAXM], fw2[MAXM];
        micro_to_macro_idx(D, K, hv->cnt, hv->idx, hv->micro, n, vm, wt, fw);
        micro_to_macro_idx(D, K, tv->cnt, tv->idx, tv->micro, n, vm, wt, fw2);
        for (int q = 0; q < M; ++q) st->mflux[(size_t)c * M + q] += area * (fw[q] + fw2[q]);
    }
    double* flux = cell_flux(o, st, c);
    for (int k = 0; k < K; ++k) {
        for (int a = 0; a < hv->cnt; ++a) flux[k * n + hv->idx[a]] += area * hv->micro[k * hv->cnt + a];
        for (int a = 0; a < tv->cnt; ++a) flux[k * n + tv->idx[a]] += area * tv->micro[k * tv->cnt + a];
    }
}

/* flux!(p4est, ka), Flux/Flux.jl:458-488 — face loop (IB phases handled by orc_ib_*) */
int orc_flux(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, double dt) {
    octx o;
    /* calc_domain_flux(DVM, Maxwellian) reads undefined variables (DVM.jl:3,12: here_weight): the combination cannot
     * run in the reference, so it is an error here too */
    if (cfg->flux_type == KAMR_FLUX_DVM)
        for (int f = 0; f < m->n_face; ++f)
            if (m->face_kind[f] == KAMR_FACE_DOMAIN && m->bc_type[m->face_there[f]] == KAMR_BC_MAXWELLIAN) return 3;
    if (octx_init(&o, cfg, m)) return 1;
    int maxn = 1;
    for (int c = 0; c < o.n_cell; ++c) if (cell_n(&o, c) > maxn) maxn = cell_n(&o, c);
    facevs hv, tv;
    hv.idx = (int*)malloc(sizeof(int) * maxn); tv.idx = (int*)malloc(sizeof(int) * maxn);
    hv.micro = (double*)malloc(sizeof(double) * (size_t)maxn * o.K);
    tv.micro = (double*)malloc(sizeof(double) * (size_t)maxn * o.K);
    /* _flux_nonib_faces! first, _flux_ib_faces! (there is a SolidNeighbor) after the wall update, Flux.jl:466-485 */
    const int sn0 = m->n_local + m->n_ghost;
    for (int pass = 0; pass < 2; ++pass)
        for (int f = 0; f < m->n_face; ++f) {
            int is_ib = m->face_kind[f] != KAMR_FACE_DOMAIN && m->face_there[f] >= sn0;
            if (is_ib != pass) continue;
            if (m->face_kind[f] == KAMR_FACE_DOMAIN) flux_domain_face(&o, st, f, dt, &hv, &tv);
            else flux_inner_face(&o, st, f, dt, &hv, &tv);
        }
    free(hv.idx); free(tv.idx); free(hv.micro); free(tv.micro);
    octx_free(&o);
    return 0;
}

/* ------------------------------------------------------------------ immersed boundary */

/* vs_extrapolate!, Boundary/Immersed_boundary.jl:98-121: adds weight[i]*(df + sdf.dx) of the source cell `src`
 * (mean over covering finer points / injection from the coarser point) to dft, laid out on the grid of cell `tgt`. */
static void vs_extrapolate(const octx* o, orc_state* st, int src, int tgt, const double* dx, const double* weight,
                           double* dft) {
    const int D = o->D, K = o->K;
    const int ns = cell_n(o, src), nt = cell_n(o, tgt);
    const int8_t* level = cell_level(o, src);
    const int8_t* levelt = cell_level(o, tgt);
    const double* df = cell_df(o, st, src);
    const double* sdf = cell_sdf(o, st, src);
    int j = 0;
    double flag = 0.0;
    for (int i = 0; i < nt; ++i) {
        if (levelt[i] == level[j]) {
            for (int k = 0; k < K; ++k) {
                double ddf = 0.0;
                for (int t = 0; t < D; ++t) ddf += sdf[(t * K + k) * ns + j] * dx[t];
                dft[k * nt + i] += (df[k * ns + j] + ddf) * weight[i];
            }
            j += 1;
        } else if (levelt[i] < level[j]) {
            while (flag != 1.0) {
                for (int k = 0; k < K; ++k) {
                    double ddf = 0.0;
                    for (int t = 0; t < D; ++t) ddf += sdf[(t * K + k) * ns + j] * dx[t];
                    dft[k * nt + i] += (df[k * ns + j] + ddf) / pow2i(D * (level[j] - levelt[i])) * weight[i];
                }
                flag += 1 / pow2i(D * (level[j] - levelt[i]));
                j += 1;
            }
            flag = 0.0;
        } else {
            for (int k = 0; k < K; ++k) {
                double ddf = 0.0;
                for (int t = 0; t < D; ++t) ddf += sdf[(t * K + k) * ns + j] * dx[t];
                dft[k * nt + i] += (df[k * ns + j] + ddf) * weight[i];
            }
            flag += 1 / pow2i(D * (levelt[i] - level[j]));
            if (flag == 1.0) { j += 1; flag = 0.0; }
        }
    }
}

/* the directional weighting shared by update_solid_cell! (:122-141) and image_df (:376-395):
 * out (K planes on the grid of cell `tgt`) = sum_i w_i(v) * extrapolate(cell_i -> point x) with
 * w_i = max(0, v.l_i/|v|)^2 normalised over i, l_i the unit vector from cell i's centre to x. */
static void directional_extrapolate(const octx* o, orc_state* st, int tgt, const double* x, int ncell,
                                    const int32_t* cells, double* out, double* weights /* n*ncell */, double* wi) {
    const int D = o->D, K = o->K, n = cell_n(o, tgt);
    const double* vm = cell_vmid(o, tgt);
    memset(out, 0, sizeof(double) * (size_t)n * K);
    for (int a = 0; a < ncell; ++a) {
        double l[MAXD], nl = 0.0;
        for (int t = 0; t < D; ++t) { l[t] = x[t] - o->m->mid[(size_t)cells[a] * D + t]; nl += l[t] * l[t]; }
        nl = sqrt(nl);
        for (int t = 0; t < D; ++t) l[t] /= nl;
        for (int j = 0; j < n; ++j) {
            double dot = 0.0, nu = 0.0;
            for (int t = 0; t < D; ++t) { dot += vm[t * n + j] * l[t]; nu += vm[t * n + j] * vm[t * n + j]; }
            double q = dot / sqrt(nu);
            q = q > 0. ? q : 0.;
            weights[(size_t)a * n + j] = q * q;
        }
    }
    for (int a = 0; a < ncell; ++a) {
        for (int j = 0; j < n; ++j) {
            double ws = 0.0;
            for (int b = 0; b < ncell; ++b) ws += weights[(size_t)b * n + j];
            wi[j] = (ws == 0.) ? 1.0 / ncell : weights[(size_t)a * n + j] / ws;
        }
        double dx[MAXD];
        for (int t = 0; t < D; ++t) dx[t] = x[t] - o->m->mid[(size_t)cells[a] * D + t];
        vs_extrapolate(o, st, cells[a], tgt, dx, wi, out);
    }
}

static int ib_maxn(const octx* o) {
    int maxn = 1;
    for (int c = 0; c < o->n_cell; ++c) if (cell_n(o, c) > maxn) maxn = cell_n(o, c);
    return maxn;
}

/* update_solid_cell!(ka), Boundary/Immersed_boundary.jl:122-141,207-213 */
int orc_ib_solid_cells(const kamr_config* cfg, const kamr_mesh* m, orc_state* st) {
    if (!m->ib) return 0;
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    const kamr_ib* ib = m->ib;
    const int D = o.D, K = o.K, M = o.M;
    int maxn = ib_maxn(&o), maxnb = 1;
    for (int s = 0; s < ib->n_solid; ++s) {
        int c = ib->solid_nb_off[s + 1] - ib->solid_nb_off[s];
        if (c > maxnb) maxnb = c;
    }
    double* weights = (double*)malloc(sizeof(double) * (size_t)maxn * maxnb);
    double* wi = (double*)malloc(sizeof(double) * (size_t)maxn);
    for (int s = 0; s < ib->n_solid; ++s) {
        int c = ib->solid_cell[s];
        int nb = ib->solid_nb_off[s + 1] - ib->solid_nb_off[s];
        directional_extrapolate(&o, st, c, m->mid + (size_t)c * D, nb, ib->solid_nb_ids + ib->solid_nb_off[s],
                                cell_df(&o, st, c), weights, wi);
        micro_to_macro_idx(D, K, cell_n(&o, c), NULL, cell_df(&o, st, c), cell_n(&o, c), cell_vmid(&o, c),
                           cell_weight(&o, c), st->w + (size_t)c * M);
        orc_get_prim(D, st->w + (size_t)c * M, cfg->gamma, st->prim + (size_t)c * M);
    }
    free(weights); free(wi);
    octx_free(&o);
    return 0;
}

/* update_solid_neighbor!(ka), Boundary/Immersed_boundary.jl:436-500 (image_df :366-395, boundary_slope! :396-435,
 * cvc_* :332-365) */
int orc_ib_solid_neighbors(const kamr_config* cfg, const kamr_mesh* m, orc_state* st) {
    if (!m->ib) return 0;
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    const kamr_ib* ib = m->ib;
    const int D = o.D, K = o.K;
    const int maxn = ib_maxn(&o);
    const int sn0 = m->n_local + m->n_ghost;
    int maxnb = 1;
    for (int s = 0; s < ib->n_sn; ++s) {
        int c = ib->sn_nb_off[s + 1] - ib->sn_nb_off[s] + 1;
        if (c > maxnb) maxnb = c;
    }
    double* weights = (double*)malloc(sizeof(double) * (size_t)maxn * maxnb);
    double* wi = (double*)malloc(sizeof(double) * (size_t)maxn);
    double* ib_df = (double*)malloc(sizeof(double) * (size_t)maxn * K);
    double* sL = (double*)malloc(sizeof(double) * (size_t)maxn * K);
    double* Mx = (double*)malloc(sizeof(double) * (size_t)maxn * K);
    double* vn = (double*)malloc(sizeof(double) * (size_t)maxn);
    double* cw = (double*)malloc(sizeof(double) * (size_t)maxn);
    double* gas_dfs = (double*)malloc(sizeof(double) * (size_t)maxn * K);
    int32_t* cells = (int32_t*)malloc(sizeof(int32_t) * (size_t)maxnb);
    for (int s = 0; s < ib->n_sn; ++s) {
        const int P = ib->sn_donor[s], S = ib->sn_solid[s], SN = sn0 + s;
        const int dir = ib->sn_faceid[s] / 2;
        const int n = cell_n(&o, P), nS = cell_n(&o, S);
        const double* vm = cell_vmid(&o, P);
        const double* wt = cell_weight(&o, P);
        const double* aux = ib->sn_aux + (size_t)s * D;
        const double* nrm = ib->sn_normal + (size_t)s * D;
        const double* pmid = m->mid + (size_t)P * D;
        const double* smid = m->mid + (size_t)S * D;
        const double* snmid = m->mid + (size_t)SN * D;
        double ib_point[MAXD];
        for (int t = 0; t < D; ++t) ib_point[t] = aux[t] + pmid[t] - smid[t];
        for (int j = 0; j < n; ++j) {
            double a = 0.0;
            for (int t = 0; t < D; ++t) a += vm[t * n + j] * nrm[t];
            vn[j] = a;
        }
        int nb = ib->sn_nb_off[s + 1] - ib->sn_nb_off[s];
        for (int a = 0; a < nb; ++a) cells[a] = ib->sn_nb_ids[ib->sn_nb_off[s] + a];
        cells[nb] = P; /* fluid_cells[end] = ps_data, :372 */
        directional_extrapolate(&o, st, P, ib_point, nb + 1, cells, ib_df, weights, wi);
        /* boundary_slope!, :396-435 */
        const double dxf = ib_point[dir] - pmid[dir], dxs = pmid[dir] - snmid[dir];
        {
            const int8_t* level = cell_level(&o, P);
            const int8_t* level_n = cell_level(&o, S);
            const double* df = cell_df(&o, st, P);
            const double* dfn = cell_df(&o, st, S);
            memset(sL, 0, sizeof(double) * (size_t)n * K);
            int index = 0;
            double flag = 0.0;
            for (int i = 0; i < n; ++i) {
                if (level[i] == level_n[index]) {
                    for (int j = 0; j < K; ++j) sL[j * n + i] = (df[j * n + i] - dfn[j * nS + index]) / dxs;
                    index += 1;
                } else if (level[i] < level_n[index]) {
                    while (flag != 1.0) {
                        for (int j = 0; j < K; ++j)
                            sL[j * n + i] += (df[j * n + i] - dfn[j * nS + index]) /
                                             pow2i(D * (level_n[index] - level[i])) / dxs;
                        flag += 1 / pow2i(D * (level_n[index] - level[i]));
                        index += 1;
                    }
                    flag = 0.0;
                } else {
                    for (int j = 0; j < K; ++j) sL[j * n + i] += (df[j * n + i] - dfn[j * nS + index]) / dxs;
                    flag += 1 / pow2i(D * (level[i] - level_n[index]));
                    if (flag == 1.0) { index += 1; flag = 0.0; }
                }
            }
            for (int j = 0; j < K; ++j)
                for (int i = 0; i < n; ++i)
                    sL[j * n + i] = minmod(sL[j * n + i], (ib_df[j * n + i] - df[j * n + i]) / dxf);
        }
        const double dxL = aux[dir] - ib_point[dir];
        double* aux_df = cell_df(&o, st, SN);
        double* ssdf = cell_sdf(&o, st, SN) + (size_t)dir * K * n;
        for (int j = 0; j < K; ++j)
            for (int i = 0; i < n; ++i) {
                double sv = sL[j * n + i];
                sv = fmin(fabs((ib_df[j * n + i] - EPS_MACH) / (sv * dxL + EPS_MACH)), 1.0) * sv; /* :452-457 */
                ssdf[j * n + i] = sv;
                aux_df[j * n + i] = ib_df[j * n + i] + sv * dxL;
            }
        /* cut velocity cells: cvc.weight = weight with the cut cells zeroed */
        const int c0 = ib->cvc_off[s], c1 = ib->cvc_off[s + 1];
        memcpy(cw, wt, sizeof(double) * (size_t)n);
        for (int q = c0; q < c1; ++q) {
            cw[ib->cvc_index[q]] = 0.0;
            for (int j = 0; j < K; ++j) gas_dfs[j * (c1 - c0) + (q - c0)] = aux_df[j * n + ib->cvc_index[q]];
        }
        double aux_prim[MAXM];
        for (int q = 0; q < D + 2; ++q) aux_prim[q] = ib->sn_bc[(size_t)s * (D + 2) + q];
        aux_prim[0] = 1.0;
        orc_discrete_maxwell(D, K, n, vm, aux_prim, cfg->K, Mx);
        double MuR = 0.0, SF = 0.0;
        for (int i = 0; i < n; ++i) { /* cvc_Mu :347-358, cvc_density :338-346 */
            double th = vn[i] >= 0 ? 1.0 : 0.0;
            MuR += cw[i] * vn[i] * Mx[i] * th;
            SF += cw[i] * vn[i] * aux_df[i] * (1.0 - th);
        }
        for (int q = c0; q < c1; ++q) {
            int i = ib->cvc_index[q];
            MuR += ib->cvc_solid_w[q] * vn[i] * Mx[i];
            SF += ib->cvc_gas_w[q] * vn[i] * gas_dfs[q - c0];
        }
        const double rho_w = -SF / MuR;
        for (int j = 0; j < K; ++j)
            for (int i = 0; i < n; ++i) Mx[j * n + i] *= rho_w;
        for (int i = 0; i < n; ++i)
            if (vn[i] >= 0)
                for (int j = 0; j < K; ++j) aux_df[j * n + i] = Mx[j * n + i];
        for (int q = c0; q < c1; ++q) { /* cvc_correction!, :360-365 */
            int i = ib->cvc_index[q];
            double gw = ib->cvc_gas_w[q], sw = ib->cvc_solid_w[q];
            for (int j = 0; j < K; ++j)
                aux_df[j * n + i] = (gw * gas_dfs[j * (c1 - c0) + (q - c0)] + sw * Mx[j * n + i]) / (gw + sw);
        }
        double* snflux = cell_flux(&o, st, SN);
        for (int j = 0; j < K; ++j)
            for (int i = 0; i < n; ++i) snflux[j * n + i] = ssdf[j * n + i] * (snmid[dir] - aux[dir]); /* :474-475 */
    }
    free(weights); free(wi); free(ib_df); free(sL); free(Mx); free(vn); free(cw); free(gas_dfs); free(cells);
    octx_free(&o);
    return 0;
}

/* ------------------------------------------------------------------ update */

/* ---- CIP_Marching: conserved I-projection, Theory/I-projection.jl ---------------------------------
 * solve_I_projection (:55-141): Newton iteration on the dual G(lambda) = <psi f exp(lambda.psi)> - W = 0, lambda0 = 0,
 * psi = (1, xi, |xi|^2/2) (build_psi_matrix :6-21), tolerance 1e-10*max(1,|W|), <= 10 iterations, stall exit
 * (:113-120), Armijo backtracking on the dual objective (:125-134), negative-f shaving first (:73-83).
 * The reference solves the (D+2)^2 Newton system with LAPACK's symmetric (Bunch-Kaufman) solver through
 * `Symmetric(J,:U) \ G` (:124); LAPACK is not restated here: Gaussian elimination with partial pivoting on the
 * symmetrised matrix.  The two differ by rounding error times cond(J) in each Newton direction, which the iteration
 * itself removes (the converged lambda is the root of G, whatever solver produced the steps). */
static void psi_of(int D, const double* vm, int n, int k, double* psi) {
    double s2 = 0.0;
    psi[0] = 1.0;
    for (int d = 0; d < D; ++d) { double x = vm[d * n + k]; psi[1 + d] = x; s2 += x * x; }
    psi[D + 1] = s2 / 2;
}
/* solves A x = b (A: M x M row-major, destroyed); returns 0 on success */
int orc_small_solve(int M, double* A, double* b, double* x) {
    for (int c = 0; c < M; ++c) {
        int piv = c;
        for (int r = c + 1; r < M; ++r)
            if (fabs(A[r * M + c]) > fabs(A[piv * M + c])) piv = r;
        if (A[piv * M + c] == 0.0) return 1;
        if (piv != c) {
            for (int q = 0; q < M; ++q) { double t = A[c * M + q]; A[c * M + q] = A[piv * M + q]; A[piv * M + q] = t; }
            double t = b[c]; b[c] = b[piv]; b[piv] = t;
        }
        for (int r = c + 1; r < M; ++r) {
            double l = A[r * M + c] / A[c * M + c];
            for (int q = c; q < M; ++q) A[r * M + q] -= l * A[c * M + q];
            b[r] -= l * b[c];
        }
    }
    for (int r = M - 1; r >= 0; --r) {
        double t = b[r];
        for (int q = r + 1; q < M; ++q) t -= A[r * M + q] * x[q];
        x[r] = t / A[r * M + r];
    }
    return 0;
}
/* dual_objective :31-47 without the - lambda.W term */
/* Test hook: the order in which the Newton sums of solve_I_projection run over the velocity points.  The algorithm's
 * exits (|G| < 1e-10 max(1,|W|), the stall test, the Armijo comparison of objectives that differ by less than their
 * rounding error near convergence) make WHICH iterate it returns depend on the rounding of these sums; running the same
 * CPU code with the points in reverse order measures how well the result is defined (tests/test_oracle_cpu.py). */
static int g_cip_reverse = 0;
void orc_set_cip_sum_order(int reverse) { g_cip_reverse = reverse; }

static double dual_sum(int D, int n, const double* vm, const double* f, const double* wt, const double* lam) {
    const int M = D + 2;
    double acc = 0.0, psi[MAXM];
    for (int kk = 0; kk < n; ++kk) {
        const int k = g_cip_reverse ? n - 1 - kk : kk;
        psi_of(D, vm, n, k, psi);
        double lp = 0.0;
        for (int i = 0; i < M; ++i) lp += lam[i] * psi[i];
        acc += wt[k] * f[k] * exp(lp);
    }
    return acc;
}
/* f is shaved in place (:73-83); lambda[D+2] out; returns the number of Newton systems solved */
int orc_solve_I_projection(int D, int n, const double* vm, double* f, const double* W, const double* wt,
                           double* lam) {
    const int M = D + 2, maxiter = 10;
    double G[MAXM], J[MAXM * MAXM], psi[MAXM];
    for (int i = 0; i < M; ++i) lam[i] = 0.0;
    double nW = 0.0;
    for (int i = 0; i < M; ++i) nW += W[i] * W[i];
    nW = sqrt(nW);
    const double tol_eff = 1e-10 * (nW > 1.0 ? nW : 1.0);
    double G_prev = INFINITY;
    int stall = 0, solves = 0;
    double fmin = f[0];
    for (int k = 1; k < n; ++k) fmin = f[k] < fmin ? f[k] : fmin;
    fmin *= 1.1;
    if (fmin < 0) {
        double fp = INFINITY;
        for (int k = 0; k < n; ++k)
            if (f[k] > 0 && f[k] < fp) fp = f[k];
        const double d = fp - fmin;
        for (int k = 0; k < n; ++k)
            if (f[k] < 0) f[k] = (f[k] - fmin) / d * fp;
    }
    for (int it = 0; it < maxiter; ++it) {
        for (int i = 0; i < M; ++i) G[i] = 0.0;
        for (int i = 0; i < M * M; ++i) J[i] = 0.0;
        for (int kk = 0; kk < n; ++kk) {
            const int k = g_cip_reverse ? n - 1 - kk : kk;
            psi_of(D, vm, n, k, psi);
            double lp = 0.0;
            for (int i = 0; i < M; ++i) lp += lam[i] * psi[i];
            const double c = wt[k] * f[k] * exp(lp);
            for (int i = 0; i < M; ++i) {
                const double ci = c * psi[i];
                G[i] += ci;
                for (int j = i; j < M; ++j) J[i * M + j] += ci * psi[j];
            }
        }
        const double Phi_sum = G[0]; /* = sum_k w f exp(lambda.psi): psi_1 = 1, same additions as dual_objective */
        double Gn = 0.0;
        for (int i = 0; i < M; ++i) { G[i] -= W[i]; Gn += G[i] * G[i]; }
        Gn = sqrt(Gn);
        if (Gn < tol_eff) return solves;
        if (Gn > 0.9 * G_prev) {
            if (++stall >= 2) return solves;
        } else stall = 0;
        G_prev = Gn;
        double A[MAXM * MAXM], b[MAXM], dl[MAXM];
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) A[i * M + j] = (j >= i) ? J[i * M + j] : J[j * M + i];
        for (int i = 0; i < M; ++i) b[i] = G[i];
        if (orc_small_solve(M, A, b, dl)) return -1;
        ++solves;
        double lW = 0.0, slope = 0.0;
        for (int i = 0; i < M; ++i) { dl[i] = -dl[i]; lW += lam[i] * W[i]; slope += G[i] * dl[i]; }
        const double Phi0 = Phi_sum - lW;
        double alpha = 1.0;
        for (int ls = 0; ls < maxiter; ++ls) {
            double lt[MAXM], ltW = 0.0;
            for (int i = 0; i < M; ++i) { lt[i] = lam[i] + alpha * dl[i]; ltW += lt[i] * W[i]; }
            if (dual_sum(D, n, vm, f, wt, lt) - ltW <= Phi0 + 1e-4 * alpha * slope) break;
            alpha *= 0.5;
        }
        for (int i = 0; i < M; ++i) lam[i] += alpha * dl[i];
    }
    return solves;
}
/* conserved_I_porjection! :144-159 (2D2F: the h-component against w with the internal energy of b removed) */
static int conserved_I_projection(int D, int K, int n, const double* vm, double* f, const double* wt,
                                  const double* w) {
    const int M = D + 2;
    double W[MAXM], lam[MAXM], psi[MAXM];
    for (int i = 0; i < M; ++i) W[i] = w[i];
    if (K > 1) {
        double e = 0.0;
        for (int k = 0; k < n; ++k) e += wt[k] * f[n + k];
        W[M - 1] -= e / 2;
    }
    if (orc_solve_I_projection(D, n, vm, f, W, wt, lam) < 0) return 1;
    for (int k = 0; k < n; ++k) {
        psi_of(D, vm, n, k, psi);
        double lp = 0.0;
        for (int i = 0; i < M; ++i) lp += lam[i] * psi[i];
        f[k] *= exp(lp);
    }
    return 0;
}

/* positivity_preserving_ib!, Boundary/Positivity.jl:1-43 (called by iterate!(CIP_Marching) on donor cells only, after
 * the macroscopic update): the slope-extrapolated correction of every SolidNeighbor face
 *   micro = (sn.flux + ndx . sn.sdf) v_n A   on the points the wall side is upwind for (rot v_dir > 0),
 * is added to w and to vs_data.flux limited by theta = min(theta_rho, theta_e) so that density and internal energy of
 * w stay positive.  `vol` = reduce(*, ds).  ORACLE ONLY: the device refuses CIP_Marching on meshes with donor cells. */
static void positivity_preserving_ib(const octx* o, orc_state* st, int c, double vol, double dt, double* w,
                                     double* vflux) {
    const kamr_mesh* m = o->m;
    const kamr_ib* ib = m->ib;
    const int D = o->D, K = o->K, M = o->M;
    if (!ib || m->bound_enc[c] == 0) return;
    const int n = cell_n(o, c);
    const double* vm = cell_vmid(o, c);
    const double* wt = cell_weight(o, c);
    const int sn0 = m->n_local + m->n_ghost;
    double* micros = (double*)calloc((size_t)n * K, sizeof(double));   /* sum over the solid faces, theta applied later */
    double* one = (double*)malloc(sizeof(double) * (size_t)n * K);
    double we[MAXM] = {0, 0, 0, 0, 0};
    for (int s = 0; s < ib->n_sn; ++s) {
        if (ib->sn_donor[s] != c) continue;
        const int SN = sn0 + s;
        const int dir = ib->sn_faceid[s] / 2;
        const double rot = (ib->sn_faceid[s] % 2 == 0) ? 1.0 : -1.0;     /* get_rot, Theory/Math.jl:2 */
        const double* snflux = cell_flux(o, st, SN);
        const double* snsdf = cell_sdf(o, st, SN);
        const double* snmid = m->mid + (size_t)SN * D;
        double fmid[MAXD], area = rot;
        for (int t = 0; t < D; ++t) {
            fmid[t] = m->mid[(size_t)c * D + t];
            if (t != dir) area *= m->ds[(size_t)c * D + t];
        }
        fmid[dir] -= 0.5 * rot * m->ds[(size_t)c * D + dir];
        memset(one, 0, sizeof(double) * (size_t)n * K);
        for (int i = 0; i < n; ++i) {
            if (!(rot * vm[dir * n + i] > 0.)) continue;
            for (int j = 0; j < K; ++j) {
                double dot = 0.0;
                for (int t = 0; t < D; ++t)
                    dot += (fmid[t] - vm[t * n + i] * dt - snmid[t]) * snsdf[(size_t)(t * K + j) * n + i];
                one[j * n + i] = (snflux[j * n + i] + dot) * vm[dir * n + i] * area;
            }
        }
        double wf[MAXM];
        micro_to_macro_idx(D, K, n, NULL, one, n, vm, wt, wf);
        for (int q = 0; q < M; ++q) we[q] += wf[q];
        for (int i = 0; i < n * K; ++i) micros[i] += one[i];
    }
    for (int q = 0; q < M; ++q) we[q] *= dt / vol;
    const double delta = 1e-3, eps = 2.220446049250313e-16;
    const double th_rho = we[0] > 0 ? 1.0 : fmin(1.0, (1 - delta) * w[0] / (fabs(we[0]) + eps));
    double rub2 = 0.0, rube = 0.0, rue2 = 0.0;
    for (int d = 1; d <= D; ++d) { rub2 += w[d] * w[d]; rube += w[d] * we[d]; rue2 += we[d] * we[d]; }
    const double eb = w[M - 1] - rub2 / (2 * w[0]);
    const double ee = we[M - 1] - rube / w[0];
    const double gam = rue2 / (2 * w[0]);
    const double th_e = fmin(1.0, 2 * (1 - delta) * eb / (sqrt(ee * ee + 4 * gam * (1 - delta) * eb) - ee + eps));
    const double th = fmin(th_rho, th_e);
    for (int q = 0; q < M; ++q) w[q] += th * we[q];
    for (int i = 0; i < n * K; ++i) vflux[i] += th * micros[i];
    free(micros); free(one);
}

/* iterate!(CAIDVM_Marching) Theory/Iterate.jl:96-130 ; iterate!(Euler) :131-162 ;
 * residual_check! Solver/Finalize.jl:5-11 */
int orc_iterate(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, double dt, int want_residual,
                double* res_out) {
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    const int D = o.D, K = o.K, M = o.M;
    double sumRes[MAXM] = {0}, sumAvg[MAXM] = {0};
    for (int c = 0; c < m->n_local; ++c) {
        if (skip_cell(&o, c)) continue;
        const int n = cell_n(&o, c);
        const double* vm = cell_vmid(&o, c);
        const double* wt = cell_weight(&o, c);
        double* f = cell_df(&o, st, c);
        double* vflux = cell_flux(&o, st, c);
        double* w = st->w + (size_t)c * M;
        double* mfl = st->mflux + (size_t)c * M;
        double area = 1.0;
        for (int d = 0; d < D; ++d) area *= m->ds[(size_t)c * D + d]; /* reduce(*, ds) */
        double prim_c[MAXM], qf[MAXD];
        if (cfg->marching == KAMR_MARCH_CAIDVM) {
            for (int q = 0; q < M; ++q) w[q] += mfl[q] * dt / area;
            orc_get_prim(D, w, cfg->gamma, prim_c);
            for (int k = 0; k < K; ++k)
                for (int i = 0; i < n; ++i) f[k * n + i] += dt / area * vflux[k * n + i];
            double w0[MAXM], prim[MAXM];
            micro_to_macro_idx(D, K, n, NULL, f, n, vm, wt, w0);
            orc_get_prim(D, w0, cfg->gamma, prim);
            double tau = orc_get_tau(D, prim_c, cfg->mu_ref, cfg->omega);
            /* f += F_c - F */
            for (int i = 0; i < n; ++i) {
                double v[MAXD], Fc[2], F[2];
                for (int t = 0; t < D; ++t) v[t] = vm[t * n + i];
                maxwell_point(D, K, v, prim_c, cfg->K, Fc);
                maxwell_point(D, K, v, prim, cfg->K, F);
                for (int k = 0; k < K; ++k) f[k * n + i] += Fc[k] - F[k];
            }
            heat_flux(D, K, n, vm, f, prim_c, wt, qf);
            for (int d = 0; d < D; ++d) st->qf[(size_t)c * D + d] = qf[d];
            for (int i = 0; i < n; ++i) {
                double v[MAXD], Fc[2], Fp[2];
                for (int t = 0; t < D; ++t) v[t] = vm[t * n + i];
                maxwell_point(D, K, v, prim_c, cfg->K, Fc);
                shakhov_point(D, K, v, Fc, prim_c, qf, cfg->Pr, cfg->K, Fp);
                for (int k = 0; k < K; ++k) {
                    Fc[k] += Fp[k];
                    double x = f[k * n + i];
                    x *= tau / (tau + dt);
                    x += dt / (tau + dt) * Fc[k];
                    f[k * n + i] = x;
                }
            }
        } else if (cfg->marching == KAMR_MARCH_EULER) {
            for (int q = 0; q < M; ++q) w[q] += mfl[q] * dt / area;
            orc_get_prim(D, w, cfg->gamma, prim_c);
            double tau = orc_get_tau(D, prim_c, cfg->mu_ref, cfg->omega);
            heat_flux(D, K, n, vm, f, prim_c, wt, qf);
            for (int d = 0; d < D; ++d) st->qf[(size_t)c * D + d] = qf[d];
            for (int i = 0; i < n; ++i) {
                double v[MAXD], F[2], Fp[2];
                for (int t = 0; t < D; ++t) v[t] = vm[t * n + i];
                maxwell_point(D, K, v, prim_c, cfg->K, F);
                shakhov_point(D, K, v, F, prim_c, qf, cfg->Pr, cfg->K, Fp);
                for (int k = 0; k < K; ++k) {
                    F[k] += Fp[k];
                    f[k * n + i] = (f[k * n + i] + dt / area * vflux[k * n + i]) * tau / (tau + dt) +
                                   dt / (tau + dt) * F[k];
                }
            }
        } else { /* iterate!(CIP_Marching), Theory/I-projection.jl:161-192 */
            for (int q = 0; q < M; ++q) w[q] += mfl[q] * dt / area;
            positivity_preserving_ib(&o, st, c, area, dt, w, vflux);   /* donor cells only */
            orc_get_prim(D, w, cfg->gamma, prim_c);
            for (int k = 0; k < K; ++k)
                for (int i = 0; i < n; ++i) f[k * n + i] += dt / area * vflux[k * n + i];
            heat_flux(D, K, n, vm, f, prim_c, wt, qf);
            for (int d = 0; d < D; ++d) st->qf[(size_t)c * D + d] = qf[d];
            if (conserved_I_projection(D, K, n, vm, f, wt, w)) { octx_free(&o); return 3; }
            double tau = orc_get_tau(D, prim_c, cfg->mu_ref, cfg->omega);
            for (int i = 0; i < n; ++i) {
                double v[MAXD], Fc[2], Fp[2];
                for (int t = 0; t < D; ++t) v[t] = vm[t * n + i];
                maxwell_point(D, K, v, prim_c, cfg->K, Fc);
                shakhov_point(D, K, v, Fc, prim_c, qf, cfg->Pr, cfg->K, Fp);
                for (int k = 0; k < K; ++k) {
                    Fc[k] += Fp[k];
                    double x = f[k * n + i];
                    x *= tau / (tau + dt);
                    x += dt / (tau + dt) * Fc[k];
                    f[k * n + i] = x;
                }
            }
        }
        if (want_residual) {
            for (int q = 0; q < M; ++q) {
                double dd = prim_c[q] - st->prim[(size_t)c * M + q];
                sumRes[q] += dd * dd;
                sumAvg[q] += fabs(prim_c[q]);
            }
        }
        for (int q = 0; q < M; ++q) { st->prim[(size_t)c * M + q] = prim_c[q]; mfl[q] = 0.0; }
        memset(vflux, 0, sizeof(double) * (size_t)n * K);
    }
    if (want_residual && res_out) {
        for (int q = 0; q < M; ++q) { res_out[q] = sumRes[q]; res_out[M + q] = sumAvg[q]; }
    }
    octx_free(&o);
    return 0;
}

/* one whole step of the solve! loop body, Solver/Solver.jl:65-67 */
int orc_step(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, double dt, int want_residual,
             double* res_out) {
    int rc = orc_slope(cfg, m, st);
    if (rc) return rc;
    rc = orc_ib_solid_cells(cfg, m, st);      /* flux!(p4est,ka): update_solid_cell!, Flux.jl:463 */
    if (rc) return rc;
    rc = orc_ib_solid_neighbors(cfg, m, st);  /* update_solid_neighbor!, Flux.jl:481 */
    if (rc) return rc;
    rc = orc_flux(cfg, m, st, dt);
    if (rc) return rc;
    return orc_iterate(cfg, m, st, dt, want_residual, res_out);
}

/* ------------------------------------------------------------------ physical-space adaptation sensor
 * update_criterion!(ka) of Physical_space/AMR.jl:256-286 with the Löhner estimator of
 * Physical_space/Criteria.jl:14-200 and the one-cell buffer apply_amr_buffer! (AMR.jl:296-341).
 * The caller has run slope! (orc_slope) first, as ps_adaptive_mesh_refinement! does (AMR.jl:1109-1111). */

/* Julia's max(a, b) returns NaN when either argument is NaN (C's fmax does not) */
static inline double jl_max(double a, double b) {
    if (a != a || b != b) return NAN;
    return a > b ? a : b;
}
#define PS_LOHNER_ABS_FLOOR 1e-4          /* Criteria.jl:2 */
#define PS_PRIMITIVE_REL_JUMP_FLOOR 1e-3  /* Criteria.jl:3 */
#define PS_VORTICITY_JUMP_FLOOR 2e-2      /* Criteria.jl:4 */

/* lohner_value, Criteria.jl:25-31 */
static double lohner_value(double left, double center, double right, double dsL, double dsR, double eps) {
    double scale = dsR * fabs(left) + (dsL + dsR) * fabs(center) + dsL * fabs(right);
    if (scale < PS_LOHNER_ABS_FLOOR * (dsL < dsR ? dsL : dsR)) return 0.0;
    double denom = dsR * fabs(left - center) + dsL * fabs(right - center) + eps * scale;
    if (denom <= 0.0) return 0.0;
    return fabs(dsR * left - (dsL + dsR) * center + dsL * right) / denom;
}
/* primitive_amplitude_ok, Criteria.jl:33-37 */
static int primitive_amplitude_ok(double left, double center, double right) {
    double jump = jl_max(fabs(left - center), fabs(right - center));
    double scale = jl_max(fabs(center), PS_LOHNER_ABS_FLOOR);
    return jump >= PS_PRIMITIVE_REL_JUMP_FLOOR * scale;
}
/* velocity_slope, Criteria.jl:39-42; sw is [dir][row] */
static inline double velocity_slope(int M, const double* sw, const double* prim, int component, int dir) {
    return (sw[dir * M + component] - prim[component] * sw[dir * M + 0]) / prim[0];
}
/* vorticity, Criteria.jl:44-53 (components 1-based there) */
static double vorticity(int D, int M, const double* sw, const double* prim) {
    if (D == 2) return velocity_slope(M, sw, prim, 1, 1) - velocity_slope(M, sw, prim, 2, 0);
    double c1 = velocity_slope(M, sw, prim, 2, 2) - velocity_slope(M, sw, prim, 3, 1);
    double c2 = velocity_slope(M, sw, prim, 3, 0) - velocity_slope(M, sw, prim, 1, 2);
    double c3 = velocity_slope(M, sw, prim, 1, 1) - velocity_slope(M, sw, prim, 2, 0);
    return sqrt(c1 * c1 + c2 * c2 + c3 * c3);
}
/* velocity_scale + vorticity_amplitude_ok, Criteria.jl:55-67 */
static int vorticity_amplitude_ok(int M, double left, double center, double right, const double* prim, double h) {
    double omega = jl_max(fabs(left), jl_max(fabs(center), fabs(right)));
    double speed2 = 0.0;
    for (int i = 1; i < M - 1; ++i) speed2 += prim[i] * prim[i];
    double lambda = jl_max(fabs(prim[M - 1]), 2.220446049250313e-16);
    double vscale = jl_max(sqrt(speed2), 1.0 / sqrt(lambda));
    return omega * h >= PS_VORTICITY_JUMP_FLOOR * vscale;
}

/* one side of update_Lohner_inner_ps!, Criteria.jl:124-160: the mean conserved state (plus the transverse shift
 * when the neighbour is coarser) turned into primitives, and the mean macro slopes.  Neighbours that are
 * SolidNeighbor pseudo-cells carry w = sw = 0 (Boundary/Immersed_boundary.jl:300-302). */
static void lohner_side(const octx* o, const orc_state* st, int c, nlist nb, int dir, int coarser, double* prim_out,
                        double* sw_out) {
    const int D = o->D, M = o->M;
    const kamr_mesh* m = o->m;
    const int n_real = m->n_local + m->n_ghost;
    double ws[MAXM];
    for (int j = 0; j < M; ++j) ws[j] = 0.0;
    for (int q = 0; q < M * D; ++q) sw_out[q] = 0.0;
    for (int k = 0; k < nb.cnt; ++k) {
        int id = nb.ids[k];
        for (int j = 0; j < M; ++j) ws[j] += id < n_real ? st->w[(size_t)id * M + j] : 0.0;
        for (int q = 0; q < M * D; ++q) sw_out[q] += id < n_real ? st->sw[(size_t)id * M * D + q] : 0.0;
    }
    if (coarser) {   /* dot(dx, sw[j, FAT[DIM-1][dir]]) over the transverse directions in ascending order */
        int id = nb.ids[0];
        for (int j = 0; j < M; ++j) {
            double acc = 0.0;
            int first = 1;
            for (int t = 0; t < D; ++t) {
                if (t == dir) continue;
                double dx = m->mid[(size_t)c * D + t] - m->mid[(size_t)id * D + t];
                double s = id < n_real ? st->sw[(size_t)id * M * D + t * M + j] : 0.0;
                if (first) { acc = dx * s; first = 0; } else acc += dx * s;
            }
            ws[j] += acc;
        }
    }
    for (int j = 0; j < M; ++j) ws[j] /= (double)nb.cnt;
    for (int q = 0; q < M * D; ++q) sw_out[q] /= (double)nb.cnt;
    orc_get_prim(D, ws, o->cfg->gamma, prim_out);
}

static double ps_sensor_of(int D, int M, const double* loh) {   /* ps_sensor, Criteria.jl:14-23 */
    double a = 0.0, b = 0.0;
    for (int d = 0; d < D; ++d) { a = jl_max(a, loh[d * M + 0]); b = jl_max(b, loh[d * M + M - 1]); }
    return jl_max(a, b);
}

/* ghost_flag: [n_ghost] the owner's "sensor > threshold" per ghost cell (lohner_flag_exchange!,
 * Parallel/Ghost.jl:939-978), NULL on a single rank.  lohner_out: [n_local][DIM][DIM+2]; sensor_out: [n_local]
 * ps_sensor after the buffer; flag_out: [n_local] the pre-buffer decision (what a rank sends for its mirrors). */
int orc_ps_criterion(const kamr_config* cfg, const kamr_mesh* m, const orc_state* st, double threshold,
                     const int32_t* ghost_flag, double* lohner_out, double* sensor_out, int32_t* flag_out) {
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    const int D = o.D, M = o.M;
    double primL[MAXM], primR[MAXM], swL[MAXM * MAXD], swR[MAXM * MAXD];
    int32_t* above = (int32_t*)calloc((size_t)m->n_local, sizeof(int32_t));
    for (int c = 0; c < m->n_local; ++c) {
        double* loh = lohner_out + (size_t)c * M * D;
        for (int q = 0; q < M * D; ++q) loh[q] = 0.0;
        if (skip_cell(&o, c)) continue;   /* AMR.jl:264-265 */
        const double* prim = st->prim + (size_t)c * M;
        const double* sw = st->sw + (size_t)c * M * D;
        for (int dir = 0; dir < D; ++dir) {
            int sL = m->nb_state[c * 2 * D + 2 * dir], sR = m->nb_state[c * 2 * D + 2 * dir + 1];
            if (sL == 0 || sR == 0) continue;   /* AMR.jl:160-255: a domain side zeroes the direction */
            double ds = m->ds[(size_t)c * D + dir];
            /* AMR.jl:5-157: same level ds, finer 0.75 ds, coarser 1.5 ds */
            double dsL = sL == 1 ? ds : (sL > 1 ? 0.75 * ds : 1.5 * ds);
            double dsR = sR == 1 ? ds : (sR > 1 ? 0.75 * ds : 1.5 * ds);
            lohner_side(&o, st, c, nb_list(&o, c, 2 * dir), dir, dsL > ds, primL, swL);
            lohner_side(&o, st, c, nb_list(&o, c, 2 * dir + 1), dir, dsR > ds, primR, swR);
            double omegaL = vorticity(D, M, swL, primL), omegaR = vorticity(D, M, swR, primR);
            double omega = vorticity(D, M, sw, prim);
            int use_vort = vorticity_amplitude_ok(M, omegaL, omega, omegaR, prim, dsL > dsR ? dsL : dsR);
            double eps_l = 0.2 * ds;
            for (int j = 0; j < M; ++j) {
                if (j == 1)
                    loh[dir * M + j] = use_vort ? lohner_value(omegaL, omega, omegaR, dsL, dsR, eps_l) : 0.0;
                else
                    loh[dir * M + j] = primitive_amplitude_ok(primL[j], prim[j], primR[j])
                                           ? lohner_value(primL[j], prim[j], primR[j], dsL, dsR, eps_l) : 0.0;
            }
        }
        above[c] = ps_sensor_of(D, M, loh) > threshold;
    }
    /* apply_amr_buffer!, AMR.jl:296-341: decide with the un-inflated sensors, then inflate */
    for (int c = 0; c < m->n_local; ++c) {
        if (flag_out) flag_out[c] = above[c];
        if (skip_cell(&o, c) || above[c]) continue;
        int flagged = 0;
        for (int f = 0; f < 2 * D && !flagged; ++f) {
            if (m->nb_state[c * 2 * D + f] == 0) continue;
            nlist nb = nb_list(&o, c, f);
            for (int k = 0; k < nb.cnt && !flagged; ++k) {
                int id = nb.ids[k];
                if (id >= m->n_local + m->n_ghost) continue;   /* SolidNeighbor: neither PsData nor GhostPsData */
                if (m->bound_enc[id] < 0) continue;
                if (id < m->n_local) flagged = above[id] == 1;   /* decide-then-apply: never a buffered cell */
                else flagged = ghost_flag ? ghost_flag[id - m->n_local] != 0 : 0;
            }
        }
        if (flagged) above[c] = 2;
    }
    for (int c = 0; c < m->n_local; ++c) {
        double* loh = lohner_out + (size_t)c * M * D;
        if (above[c] == 2)
            for (int q = 0; q < M * D; ++q) loh[q] = 2.0 * threshold;
        if (sensor_out) sensor_out[c] = ps_sensor_of(D, M, loh);
    }
    free(above);
    octx_free(&o);
    return 0;
}

/* ------------------------------------------------------------------ velocity-space adaptation inputs
 * vs_resolution (Velocity_space/AMR.jl:139-166), the refine_flags of vs_refine! (:26-69) and the coarsen_ok of
 * vs_coarsen! (:73-115) with the criteria of Velocity_space/Criteria.jl and the face-neighbour search of
 * Velocity_space/Neighbor.jl.  The neighbour search here is a different algorithm with the same answer: instead of
 * the reference's sorted Morton keys + predecessor search, the leaf owning a finest-level lattice point is found by
 * looking its ancestors' corners up, level by level, in a sorted table of (level, corner) keys. */

typedef struct { uint64_t key; int leaf; } vkey;
static int vkey_cmp(const void* a, const void* b) {
    uint64_t x = ((const vkey*)a)->key, y = ((const vkey*)b)->key;
    return x < y ? -1 : (x > y ? 1 : 0);
}
/* 4 bits of level, 20 bits per finest-level lattice coordinate */
static inline uint64_t vkey_pack(int D, int level, const long long* g) {
    uint64_t k = (uint64_t)level;
    for (int d = 0; d < D; ++d) k = (k << 20) | (uint64_t)g[d];
    return k;
}
typedef struct {
    int D, n, maxlevel;
    double vmin[MAXD], h_fine[MAXD];
    long long gmax[MAXD];
    vkey* keys;
} vindex;

static int vindex_build(vindex* ix, int D, int n, const double* vmid, const int8_t* level, const kamr_vs_adapt* par) {
    ix->D = D; ix->n = n; ix->maxlevel = par->maxlevel;
    if (par->maxlevel > 15) return 1;
    for (int d = 0; d < D; ++d) {
        double ds0 = (par->vmax[d] - par->vmin[d]) / par->trees[d];   /* AMR.jl:29-30 */
        ix->vmin[d] = par->vmin[d];
        ix->h_fine[d] = ds0 / pow2i(par->maxlevel);                   /* Neighbor.jl:83 */
        ix->gmax[d] = (long long)par->trees[d] << par->maxlevel;
        if (ix->gmax[d] >= (1 << 20)) return 1;
    }
    ix->keys = (vkey*)malloc(sizeof(vkey) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
        long long g[MAXD];
        int L = level[i];
        for (int d = 0; d < D; ++d) {   /* _corner_index, Neighbor.jl:63-66 */
            double cell = ix->h_fine[d] * (double)(1 << (par->maxlevel - L));
            g[d] = llround((vmid[(size_t)d * n + i] - 0.5 * cell - ix->vmin[d]) / ix->h_fine[d]);
        }
        ix->keys[i].key = vkey_pack(D, L, g);
        ix->keys[i].leaf = i;
    }
    qsort(ix->keys, (size_t)n, sizeof(vkey), vkey_cmp);
    return 0;
}
/* leaf owning the finest-level lattice point g, -1 if none */
static int vindex_locate(const vindex* ix, const long long* g) {
    for (int L = ix->maxlevel; L >= 0; --L) {
        long long c[MAXD];
        int sh = ix->maxlevel - L;
        for (int d = 0; d < ix->D; ++d) c[d] = (g[d] >> sh) << sh;
        vkey probe = { vkey_pack(ix->D, L, c), 0 };
        const vkey* hit = (const vkey*)bsearch(&probe, ix->keys, (size_t)ix->n, sizeof(vkey), vkey_cmp);
        if (hit) return hit->leaf;
    }
    return -1;
}
/* vs_face_neighbor, Neighbor.jl:181-204: probe half a finest cell across the face, on the cell's centre line */
static int vindex_face_neighbor(const vindex* ix, const double* vmid, const int8_t* level, int i, int dim, int dir) {
    long long g[MAXD];
    int L = level[i];
    double cell = ix->h_fine[dim] * (double)(1 << (ix->maxlevel - L));
    for (int d = 0; d < ix->D; ++d) {
        double coord = vmid[(size_t)d * ix->n + i];
        if (d == dim) coord += dir * (0.5 * cell + 0.5 * ix->h_fine[d]);
        double q = floor((coord - ix->vmin[d]) / ix->h_fine[d]);
        if (q < 0 || q >= (double)ix->gmax[d]) return -1;
        g[d] = (long long)q;
    }
    return vindex_locate(ix, g);
}

/* exported for the tests: face neighbours [n][DIM][2] (-1: velocity-domain boundary) of one velocity grid */
int orc_vs_face_neighbors(int D, int n, const double* vmid, const int8_t* level, const kamr_vs_adapt* par,
                          int32_t* out) {
    vindex ix;
    if (vindex_build(&ix, D, n, vmid, level, par)) return 1;
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < D; ++d) {
            out[((size_t)i * D + d) * 2 + 0] = vindex_face_neighbor(&ix, vmid, level, i, d, -1);
            out[((size_t)i * D + d) * 2 + 1] = vindex_face_neighbor(&ix, vmid, level, i, d, +1);
        }
    free(ix.keys);
    return 0;
}

/* vs_resolution(ps_data, kinfo) maximised over the local fluid cells, AMR.jl:139-166 */
int orc_vs_resolution(const kamr_config* cfg, const kamr_mesh* m, const orc_state* st, const kamr_vs_adapt* par,
                      double* out) {
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    const int D = o.D, K = o.K, M = o.M;
    double du = 1.0, nt = 1.0;
    for (int d = 0; d < D; ++d) { du *= par->vmax[d] - par->vmin[d]; nt *= (double)par->trees[d]; }
    double weight = du / nt / pow2i(D * par->maxlevel);
    double dres = 0.0, eres = 0.0;
    for (int c = 0; c < m->n_local; ++c) {
        if (skip_cell(&o, c)) continue;
        int n = cell_n(&o, c);
        const double* df = st->df + o.vs_off[c] * K;
        const double* v = cell_vmid(&o, c);
        const double* U = st->prim + (size_t)c * M + 1;
        double dmax = -INFINITY, emax = -INFINITY;
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k < K; ++k) dmax = jl_max(dmax, df[(size_t)k * n + i]);   /* maximum(vs_data.df) */
            double c2 = 0.0;
            for (int d = 0; d < D; ++d) { double t = U[d] - v[(size_t)d * n + i]; c2 += t * t; }
            double e = K == 1 ? df[i] * c2 : df[i] * c2 + df[(size_t)n + i];
            emax = jl_max(emax, e);
        }
        dres = jl_max(dmax * weight, dres);
        eres = jl_max(0.5 * emax * weight, eres);
    }
    out[0] = dres; out[1] = eres;
    octx_free(&o);
    return 0;
}

#define VS_LOHNER_EPS_FLT 1e-2        /* Criteria.jl:213 */
#define VS_LOHNER_EPS_ABS 1e-3        /* Criteria.jl:215 */
#define VS_LOHNER_COARSEN_RATIO 0.3   /* Criteria.jl:217 */

/* _lohner_ratio, Criteria.jl:222-228 */
static double vs_lohner_ratio(double L, double C, double R, double dsL, double dsR, double scale) {
    double num = fabs(dsR * L - (dsL + dsR) * C + dsL * R);
    double den = dsR * fabs(L - C) + dsL * fabs(R - C) +
                 VS_LOHNER_EPS_FLT * (dsR * fabs(L) + (dsL + dsR) * fabs(C) + dsL * fabs(R)) +
                 VS_LOHNER_EPS_ABS * scale * (dsL + dsR);
    return den > 0 ? num / den : 0.0;
}

int orc_vs_criterion(const kamr_config* cfg, const kamr_mesh* m, const orc_state* st, const kamr_vs_adapt* par,
                     uint8_t* refine_flag, uint8_t* coarsen_ok) {
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    const int D = o.D, K = o.K, M = o.M;
    const int lohner = par->mode == 0;
    int rc = 0;
    /* one neighbour table per velocity grid */
    int32_t** nbt = (int32_t**)calloc((size_t)m->n_grid, sizeof(int32_t*));
    for (int c = 0; c < m->n_local && !rc; ++c) {
        const int n = cell_n(&o, c), grid = m->cell_grid[c];
        const double* v = cell_vmid(&o, c);
        const int8_t* lev = cell_level(&o, c);
        const double* wt = cell_weight(&o, c);
        const double* df = st->df + o.vs_off[c] * K;
        const double* sdf = st->sdf + o.vs_off[c] * K * D;
        const double* w = st->w + (size_t)c * M;
        const double* U = st->prim + (size_t)c * M + 1;
        const double* ds = m->ds + (size_t)c * D;
        uint8_t* rf = refine_flag ? refine_flag + o.vs_off[c] : NULL;
        uint8_t* co = coarsen_ok ? coarsen_ok + o.vs_off[c] : NULL;
        double s1 = 0.0, s2 = 0.0;
        if (lohner) {
            if (!nbt[grid]) {
                nbt[grid] = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1) * D * 2);
                if (orc_vs_face_neighbors(D, n, v, lev, par, nbt[grid])) { rc = 2; break; }
            }
            for (int i = 0; i < n; ++i) {   /* vs_lohner_scales, Criteria.jl:238-248 */
                double a1 = fabs(df[i]); if (a1 > s1) s1 = a1;
                if (K == 2) { double a2 = fabs(df[(size_t)n + i]); if (a2 > s2) s2 = a2; }
            }
        }
        double U2 = 0.0;
        for (int d = 0; d < D; ++d) U2 += U[d] * U[d];
        const double eden = w[M - 1] - 0.5 * w[0] * U2;   /* w[end] - 0.5 w[1] sum(U.^2) */
        const double two_d = pow2i(D);
        for (int i = 0; i < n; ++i) {
            double cdf[2] = {0.0, 0.0};
            for (int k = 0; k < K; ++k) {   /* _criterion_cell!, AMR.jl:8-21 */
                double mx = 0.0;
                for (int d = 0; d < D; ++d) {
                    double a = fabs(sdf[((size_t)d * K + k) * n + i] * ds[d]);
                    if (a > mx) mx = a;
                }
                cdf[k] = df[(size_t)k * n + i] + mx;
            }
            double S = 0.0;
            for (int d = 0; d < D; ++d) { double t = U[d] - v[(size_t)d * n + i]; S += t * t; }
            const double wgt = wt[i];
            /* local_contribution_refine_flag, Criteria.jl:19-26 */
            double e_ref = K == 2 ? fabs(0.5 * (S * cdf[0] + cdf[1]) * wgt) : fabs(0.5 * S * cdf[0] * wgt);
            int local_refine = jl_max(e_ref / eden, cdf[0] * wgt / w[0]) > par->coeff_local;
            /* local_contribution_coarsen_flag (vector form), Criteria.jl:41-48 */
            double e_co = K == 2 ? 0.5 * (S * cdf[0] + cdf[1]) * wgt : 0.5 * (S * cdf[0]) * wgt;
            int local_coarsen = jl_max(e_co / eden, cdf[0] * wgt / w[0]) < par->coeff_local / two_d;
            int base_refine, ok;
            if (lohner) {
                /* vs_lohner_indicator, Criteria.jl:259-287 (on df, not the criterion distribution) */
                double eta = 0.0;
                const int32_t* nb = nbt[grid] + (size_t)i * D * 2;
                for (int d = 0; d < D; ++d) {
                    double hfine = ((par->vmax[d] - par->vmin[d]) / par->trees[d]) / pow2i(par->maxlevel);
                    double hi = hfine * (double)(1 << (par->maxlevel - lev[i]));
                    int Ln = nb[d * 2], Rn = nb[d * 2 + 1];
                    double dsL = Ln < 0 ? hi : 0.5 * (hi + hfine * (double)(1 << (par->maxlevel - lev[Ln])));
                    double dsR = Rn < 0 ? hi : 0.5 * (hi + hfine * (double)(1 << (par->maxlevel - lev[Rn])));
                    for (int k = 0; k < K; ++k) {
                        double fL = Ln < 0 ? 0.0 : df[(size_t)k * n + Ln], fR = Rn < 0 ? 0.0 : df[(size_t)k * n + Rn];
                        eta = jl_max(eta, vs_lohner_ratio(fL, df[(size_t)k * n + i], fR, dsL, dsR, k == 0 ? s1 : s2));
                    }
                }
                base_refine = eta > par->coeff_lohner || local_refine;
                ok = eta < VS_LOHNER_COARSEN_RATIO * par->coeff_lohner && local_coarsen;
            } else {
                /* global_contribution_{refine,coarsen}_flag, Criteria.jl:55-84 */
                double e_gl = K == 2 ? 0.5 * (S * cdf[0] + cdf[1]) * wgt : 0.5 * (S * cdf[0]) * wgt;
                int global_refine = cdf[0] * wgt > par->coeff_global * par->vr_density ||
                                    e_gl > par->vr_energy * par->coeff_global;
                int global_coarsen = cdf[0] * wgt < par->coeff_global * par->vr_density / two_d &&
                                     e_gl < par->vr_energy * par->coeff_global / two_d;
                base_refine = local_refine || global_refine;
                ok = local_coarsen && global_coarsen;
            }
            if (rf) rf[i] = (uint8_t)(lev[i] < par->maxlevel && base_refine);
            if (co) co[i] = (uint8_t)ok;
        }
    }
    for (int g = 0; g < m->n_grid; ++g) free(nbt[g]);
    free(nbt);
    octx_free(&o);
    return rc;
}

/* vs_conserved_correction!, Velocity_space/AMR.jl:120-133: conserved_I_porjection!(vs_data, ps_data.w) on the listed
 * local cells (the ones a velocity-space adaptation pass regridded); solid cells are skipped as there */
int orc_project_cells(const kamr_config* cfg, const kamr_mesh* m, orc_state* st, int n_list, const int32_t* list) {
    octx o;
    if (octx_init(&o, cfg, m)) return 1;
    int rc = 0;
    for (int q = 0; q < n_list && !rc; ++q) {
        int c = list[q];
        if (c < 0 || c >= m->n_local) { rc = 2; break; }
        if (skip_cell(&o, c)) continue;
        rc = conserved_I_projection(o.D, o.K, cell_n(&o, c), cell_vmid(&o, c), cell_df(&o, st, c), cell_weight(&o, c),
                                    st->w + (size_t)c * o.M);
    }
    octx_free(&o);
    return rc;
}
