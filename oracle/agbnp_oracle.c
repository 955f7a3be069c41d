/* ORACLE / TEST INFRASTRUCTURE ONLY -- see agbnp_oracle.h for scope and parity status.
 *
 * Plain-C restatement (double arithmetic, same expression order) of
 *   gaussvol/gaussvol.cpp                                   GaussVol overlap tree
 *   openmmapi/src/AGBNPUtils.cpp (+ OpenMM SplineFitter)    I4 / Q4 lookup tables
 *   platforms/reference/src/ReferenceAGBNPKernels.cpp       executeGVolSA / executeAGBNP1
 * Each function cites the reference lines it follows.  Compile with -ffp-contract=off (see Makefile).
 */
#include "agbnp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants: float literals promoted to double exactly as the reference TUs see them ---- */
#define KFC ((double)2.2269859253f)               /* gaussvol.h:46 */
#define PFC ((double)2.5f)                        /* gaussvol.h:47 */
#define MIN_GVOL ((double)FLT_MIN)                /* gaussvol.h:52 */
#define MAX_ORDER 8                               /* gaussvol.h:55 */
#define ANG (0.1f)                                /* gaussvol.h:58 */
#define ANG3 (0.001f)                             /* gaussvol.h:59 */
#define VOLMINA ((double)(0.01f*ANG3))            /* gaussvol.h:62 (float product, then promoted) */
#define VOLMINB ((double)(0.1f*ANG3))             /* gaussvol.h:63 */
#define AGBNP_RADIUS_INCREMENT ((double)(0.5f*ANG)) /* AGBNPForce.h:25 */
#define AGBNP_HB_RADIUS (1.4*(double)ANG)         /* AGBNPForce.h:33 (double x float) */
#define AGBNP_I4LOOKUP_MAXA (2.0)                 /* AGBNPUtils.h:124 */
#define AGBNP_I4LOOKUP_NA (16)                    /* AGBNPUtils.h:126 */
#define AGBNP_RADIUS_PRECISION (10000)            /* AGBNPUtils.h:155 */

static char g_err[512] = "";
const char* agbnp_oracle_last_error(void) { return g_err; }
static void set_err(const char* m) { snprintf(g_err, sizeof g_err, "%s", m); }

/* ------------------------------------------------------------------------------------------------------------- */
/* GaussVol                                                                                                        */
/* ------------------------------------------------------------------------------------------------------------- */
typedef struct { double v, a, c[3]; } gaussian_vca;              /* gaussvol.h:66-71 */

typedef struct {                                                 /* gaussvol.h:96-112 */
    int level;
    gaussian_vca g;
    double volume, dvv1, dv1[3], gamma1i, self_volume, sfp;
    int atom, parent_index, children_startindex, children_count;
} goverlap;

typedef struct {
    int natoms;
    goverlap* ov; int size, cap;
    long n_eval2, n_eval3;      /* ogauss evaluations at level 2 / deeper (work counters, not in the reference) */
} gtree;

static void tree_push(gtree* t, const goverlap* o) {
    if (t->size == t->cap) {
        t->cap = t->cap ? 2*t->cap : 1024;
        t->ov = (goverlap*) realloc(t->ov, sizeof(goverlap)*(size_t)t->cap);
    }
    t->ov[t->size++] = *o;
}

/* gaussvol.cpp:18-41 */
static double pol_switchfunc(double gvol, double volmina, double volminb, double* sp) {
    double swf = 0.0, swfp = 1.0, swd, swu, swu2, swu3, s;
    if (gvol > volminb) { swf = 1.0; swfp = 0.0; }
    else if (gvol < volmina) { swf = 0.0; swfp = 0.0; }
    swd = 1.0/(volminb - volmina);
    swu = (gvol - volmina)*swd;
    swu2 = swu*swu;
    swu3 = swu*swu2;
    s = swf + swfp*swu3*(10.0 - 15.0*swu + 6.0*swu2);
    *sp = swfp*swd*30.0*swu2*(1.0 - 2.0*swu + swu2);
    return s;
}

/* gaussvol.cpp:60-93 */
static double ogauss_alpha(const gaussian_vca* g1, const gaussian_vca* g2, gaussian_vca* g12,
                           double* dVdr, double* dVdV, double* sfp) {
    double dist[3], d2, a12, deltai, df, ef, gvol, dgvol, dgvolv, s, sp;
    int k;
    for (k = 0; k < 3; k++) dist[k] = g2->c[k] - g1->c[k];
    d2 = dist[0]*dist[0] + dist[1]*dist[1] + dist[2]*dist[2];
    a12 = g1->a + g2->a;
    deltai = 1.0/a12;
    df = (g1->a)*(g2->a)*deltai;
    ef = exp(-df*d2);
    gvol = ((g1->v*g2->v)/pow(M_PI/df, 1.5))*ef;
    dgvol = -2.0*df*gvol;
    dgvolv = g1->v > 0 ? gvol/g1->v : 0.0;
    for (k = 0; k < 3; k++) g12->c[k] = ((g1->c[k]*g1->a) + (g2->c[k]*g2->a))*deltai;
    g12->a = a12;
    g12->v = gvol;
    s = pol_switchfunc(gvol, VOLMINA, VOLMINB, &sp);
    *sfp = sp*gvol + s;
    *dVdr = dgvol;
    *dVdV = dgvolv;
    return s*gvol;
}

/* gaussvol.cpp:103-151 */
static void init_overlap_tree(gtree* t, const double* pos, const double* radius, const double* volume,
                              const double* gamma, const int* ishydrogen) {
    goverlap o;
    int iat, k;
    memset(&o, 0, sizeof o);
    t->size = 0;
    o.level = 0; o.volume = 0; o.dvv1 = 0.0; o.self_volume = 0; o.sfp = 1.0; o.gamma1i = 0.0;
    o.parent_index = -1; o.atom = -1; o.children_startindex = 1; o.children_count = t->natoms;
    tree_push(t, &o);
    for (iat = 0; iat < t->natoms; iat++) {
        double a = KFC/(radius[iat]*radius[iat]);
        double vol = ishydrogen[iat] > 0 ? 0.0 : volume[iat];
        o.level = 1;
        o.g.v = vol; o.g.a = a;
        for (k = 0; k < 3; k++) { o.g.c[k] = pos[3*iat+k]; o.dv1[k] = 0.0; }
        o.volume = vol; o.dvv1 = 1.0; o.self_volume = 0.0; o.sfp = 1.0; o.gamma1i = gamma[iat];
        o.parent_index = 0; o.atom = iat; o.children_startindex = -1; o.children_count = -1;
        tree_push(t, &o);
    }
}

/* gaussvol.cpp:97-100: larger volume first.  The reference uses std::sort (introsort, unstable on ties); ties are
 * exact-equality of doubles and do not occur on the fixtures.  A stable merge sort is used here so that the order
 * under ties is defined (input order = sibling order); the comparison itself is the reference's. */
static void sort_children(goverlap* c, int n, goverlap* tmp) {
    int w, i;
    for (w = 1; w < n; w *= 2) {
        for (i = 0; i < n; i += 2*w) {
            int l = i, m = i+w < n ? i+w : n, r = i+2*w < n ? i+2*w : n, a = l, b = m, k = l;
            while (a < m && b < r) tmp[k++] = (c[b].volume > c[a].volume) ? c[b++] : c[a++];
            while (a < m) tmp[k++] = c[a++];
            while (b < r) tmp[k++] = c[b++];
        }
        memcpy(c, tmp, sizeof(goverlap)*(size_t)n);
    }
}

/* exact-safe prefilter (not in the reference): a candidate whose unswitched volume is certainly below VOLMINA has
 * switched volume exactly 0 and is rejected by the `gvol > MIN_GVOL` test (gaussvol.cpp:233) anyway. */
static int certainly_no_overlap(const gaussian_vca* g1, const gaussian_vca* g2) {
    double d2 = 0, df, pref, lim;
    int k;
    if (!(g1->v > 0) || !(g2->v > 0)) return 1;   /* hydrogens: gvol == 0 exactly */
    for (k = 0; k < 3; k++) { double d = g2->c[k]-g1->c[k]; d2 += d*d; }
    df = g1->a*g2->a/(g1->a+g2->a);
    pref = g1->v*g2->v*pow(df/M_PI, 1.5);
    if (pref <= VOLMINA) return 1;
    lim = log(pref/VOLMINA)/df;
    return d2 > 1.02*lim + 1e-6;
}

/* gaussvol.cpp:197-250 (compute_children) + :154-192 (add_children) + :376-387 (recursion) */
static void compute_andadd_children_r(gtree* t, int root_index) {
    int parent_index, sibling_start, sibling_count, slotj, n = 0, cap = 64, start, i, k;
    goverlap *children, *tmp;
    int root_level;
    {
        goverlap* root = &t->ov[root_index];
        parent_index = root->parent_index;
        if (parent_index < 0) return;
        if (root->level >= MAX_ORDER) return;
        sibling_start = t->ov[parent_index].children_startindex;
        sibling_count = t->ov[parent_index].children_count;
        if (sibling_start < 0 || sibling_count < 0) return;
        root_level = root->level;
    }
    children = (goverlap*) malloc(sizeof(goverlap)*(size_t)cap);
    for (slotj = root_index+1; slotj < sibling_start+sibling_count; slotj++) {
        gaussian_vca g12;
        double gvol, dVdr, dVdV, sfp;
        const goverlap* root = &t->ov[root_index];
        int atom2 = t->ov[slotj].atom;
        const gaussian_vca* g1 = &root->g;
        const gaussian_vca* g2 = &t->ov[atom2+1].g;
        if (certainly_no_overlap(g1, g2)) continue;
        if (root_level == 1) t->n_eval2++; else t->n_eval3++;
        gvol = ogauss_alpha(g1, g2, &g12, &dVdr, &dVdV, &sfp);
        if (gvol > MIN_GVOL) {
            goverlap ov;
            memset(&ov, 0, sizeof ov);
            ov.g = g12;
            ov.volume = gvol;
            ov.self_volume = 0;
            ov.atom = atom2;
            for (k = 0; k < 3; k++) ov.dv1[k] = (g2->c[k] - g1->c[k])*(-dVdr);
            ov.dvv1 = dVdV;
            ov.sfp = sfp;
            ov.gamma1i = root->gamma1i + t->ov[atom2+1].gamma1i;
            if (n == cap) { cap *= 2; children = (goverlap*) realloc(children, sizeof(goverlap)*(size_t)cap); }
            children[n++] = ov;
        }
    }
    if (n > 0) {
        tmp = (goverlap*) malloc(sizeof(goverlap)*(size_t)n);
        sort_children(children, n, tmp);
        free(tmp);
        start = t->size;
        t->ov[root_index].children_startindex = start;
        t->ov[root_index].children_count = n;
        for (i = 0; i < n; i++) {
            children[i].level = root_level+1;
            children[i].parent_index = root_index;
            children[i].children_startindex = -1;
            children[i].children_count = -1;
            tree_push(t, &children[i]);
        }
        free(children);
        for (i = start; i < start+n; i++) compute_andadd_children_r(t, i);
    } else {
        free(children);
    }
}

/* gaussvol.cpp:389-397 */
static void compute_overlap_tree_r(gtree* t, const double* pos, const double* radius, const double* volume,
                                   const double* gamma, const int* ishydrogen) {
    int slot;
    t->n_eval2 = t->n_eval3 = 0;
    init_overlap_tree(t, pos, radius, volume, gamma, ishydrogen);
    for (slot = 1; slot <= t->natoms; slot++) compute_andadd_children_r(t, slot);
}

/* gaussvol.cpp:254-287 */
static void rescan_r(gtree* t, int slot) {
    goverlap* ov = &t->ov[slot];
    int parent_index = ov->parent_index, k, c;
    if (parent_index > 0) {
        gaussian_vca g12;
        double dVdr, dVdV, sfp, gvol;
        int atom = ov->atom;
        const gaussian_vca* g1 = &t->ov[parent_index].g;
        const gaussian_vca* g2 = &t->ov[atom+1].g;
        gvol = ogauss_alpha(g1, g2, &g12, &dVdr, &dVdV, &sfp);
        ov->g = g12;
        ov->volume = gvol;
        for (k = 0; k < 3; k++) ov->dv1[k] = (g2->c[k] - g1->c[k])*(-dVdr);
        ov->dvv1 = dVdV;
        ov->sfp = sfp;
        ov->gamma1i = t->ov[parent_index].gamma1i + t->ov[atom+1].gamma1i;
    }
    for (c = ov->children_startindex; c < ov->children_startindex+ov->children_count; c++) rescan_r(t, c);
}

/* gaussvol.cpp:290-327 */
static void rescan_tree_v(gtree* t, const double* pos, const double* radius, const double* volume,
                          const double* gamma, const int* ishydrogen) {
    int iat, k;
    goverlap* ov = &t->ov[0];
    ov->level = 0; ov->volume = 0; ov->dv1[0] = ov->dv1[1] = ov->dv1[2] = 0; ov->dvv1 = 0.0; ov->self_volume = 0;
    ov->sfp = 1.0; ov->gamma1i = 0.0;
    for (iat = 0; iat < t->natoms; iat++) {
        double a = KFC/(radius[iat]*radius[iat]);
        double vol = ishydrogen[iat] > 0 ? 0.0 : volume[iat];
        ov = &t->ov[iat+1];
        ov->level = 1; ov->g.v = vol; ov->g.a = a;
        for (k = 0; k < 3; k++) { ov->g.c[k] = pos[3*iat+k]; ov->dv1[k] = 0; }
        ov->volume = vol; ov->dvv1 = 1.0; ov->self_volume = 0.0; ov->sfp = 1.0; ov->gamma1i = gamma[iat];
    }
    rescan_r(t, 0);
}

/* gaussvol.cpp:330-351 */
static void rescan_gamma_r(gtree* t, int slot) {
    goverlap* ov = &t->ov[slot];
    int parent_index = ov->parent_index, c;
    if (parent_index > 0) ov->gamma1i = t->ov[parent_index].gamma1i + t->ov[ov->atom+1].gamma1i;
    for (c = ov->children_startindex; c < ov->children_startindex+ov->children_count; c++) rescan_gamma_r(t, c);
}

/* gaussvol.cpp:356-372 */
static void rescan_tree_g(gtree* t, const double* gamma) {
    int iat;
    t->ov[0].gamma1i = 0.0;
    for (iat = 0; iat < t->natoms; iat++) t->ov[iat+1].gamma1i = gamma[iat];
    rescan_gamma_r(t, 0);
}

typedef struct { double psi, f, p[3]; } acc3;

/* gaussvol.cpp:400-487 */
static void compute_volume_underslot2_r(gtree* t, int slot, acc3* fv, acc3* sv, acc3* en,
                                        double* dr, double* dv, double* free_volume, double* self_volume) {
    const goverlap* ov = &t->ov[slot];
    double cf = ov->level % 2 == 0 ? -1.0 : 1.0;
    double volcoeff = ov->level > 0 ? cf : 0;
    double volcoeffp = ov->level > 0 ? volcoeff/(double)ov->level : 0;
    int atom = ov->atom, k, sloti;
    double ai = t->ov[atom+1].g.a;   /* atom == -1 for the root reads slot 0; unused there (level 0) */
    double a1i = ov->g.a;
    double a1 = a1i - ai;
    double c2;

    fv->psi = volcoeff*ov->volume; fv->f = volcoeff*ov->sfp;
    sv->psi = volcoeffp*ov->volume; sv->f = volcoeffp*ov->sfp;
    en->psi = volcoeffp*ov->gamma1i*ov->volume; en->f = volcoeffp*ov->sfp*ov->gamma1i;
    for (k = 0; k < 3; k++) fv->p[k] = sv->p[k] = en->p[k] = 0.0;

    if (ov->children_startindex >= 0) {
        for (sloti = ov->children_startindex; sloti < ov->children_startindex+ov->children_count; sloti++) {
            acc3 fvt, svt, ent;
            compute_volume_underslot2_r(t, sloti, &fvt, &svt, &ent, dr, dv, free_volume, self_volume);
            ov = &t->ov[slot];
            fv->psi += fvt.psi; fv->f += fvt.f;
            sv->psi += svt.psi; sv->f += svt.f;
            en->psi += ent.psi; en->f += ent.f;
            for (k = 0; k < 3; k++) { fv->p[k] += fvt.p[k]; sv->p[k] += svt.p[k]; en->p[k] += ent.p[k]; }
        }
    }
    if (ov->level > 0) {
        free_volume[atom] += fv->psi;
        self_volume[atom] += sv->psi;
        c2 = ai/a1i;
        for (k = 0; k < 3; k++) dr[3*atom+k] += (-ov->dv1[k])*en->f + en->p[k]*c2;
        dv[atom] += ov->g.v*en->f;
        c2 = a1/a1i;
        for (k = 0; k < 3; k++) {
            fv->p[k] = ov->dv1[k]*fv->f + fv->p[k]*c2;
            sv->p[k] = ov->dv1[k]*sv->f + sv->p[k]*c2;
            en->p[k] = ov->dv1[k]*en->f + en->p[k]*c2;
        }
        fv->f = ov->dvv1*fv->f;
        sv->f = ov->dvv1*sv->f;
        en->f = ov->dvv1*en->f;
    }
}

/* gaussvol.cpp:490-519 + GaussVol::compute_volume :589-606 */
static void compute_volume(gtree* t, const double* volumes, double* volume, double* energy, double* force,
                           double* gradV, double* free_volume, double* self_volume) {
    acc3 fv, sv, en;
    int i, n = t->natoms;
    for (i = 0; i < 3*n; i++) force[i] = 0.0;
    for (i = 0; i < n; i++) gradV[i] = free_volume[i] = self_volume[i] = 0.0;
    compute_volume_underslot2_r(t, 0, &fv, &sv, &en, force, gradV, free_volume, self_volume);
    *volume = fv.psi;
    *energy = en.psi;
    for (i = 0; i < 3*n; i++) force[i] = -force[i];
    for (i = 0; i < n; i++) if (volumes[i] > 0) gradV[i] = gradV[i]/volumes[i];
}

/* ------------------------------------------------------------------------------------------------------------- */
/* natural cubic spline: OpenMM SplineFitter (third-party, not under /root/reference) as called at              */
/* AGBNPUtils.h:104,112,115.  Same arithmetic as oracle/shim/openmm/internal/SplineFitter.h.                     */
/* ------------------------------------------------------------------------------------------------------------- */
static void create_natural_spline(int n, const double* x, const double* y, double* y2) {
    double *lo, *di, *up, *rhs, *g, beta;
    int i;
    for (i = 0; i < n; i++) y2[i] = 0.0;
    if (n == 2) return;
    lo = (double*) calloc((size_t)5*n, sizeof(double)); di = lo+n; up = di+n; rhs = up+n; g = rhs+n;
    for (i = 0; i < n; i++) di[i] = 1.0;
    for (i = 1; i < n-1; i++) {
        lo[i] = x[i]-x[i-1];
        di[i] = 2.0*(x[i+1]-x[i-1]);
        up[i] = x[i+1]-x[i];
        rhs[i] = 6.0*((y[i+1]-y[i])/(x[i+1]-x[i]) - (y[i]-y[i-1])/(x[i]-x[i-1]));
    }
    y2[0] = rhs[0]/di[0];
    beta = di[0];
    for (i = 1; i < n; i++) {
        g[i] = up[i-1]/beta;
        beta = di[i]-lo[i]*g[i];
        y2[i] = (rhs[i]-lo[i]*y2[i-1])/beta;
    }
    for (i = n-2; i >= 0; i--) y2[i] -= g[i+1]*y2[i+1];
    free(lo);
}

static void spline_locate(int n, const double* x, double t, int* lower, int* upper) {
    *lower = 0; *upper = n-1;
    while (*upper-*lower > 1) {
        int middle = (*upper+*lower)/2;
        if (x[middle] > t) *upper = middle; else *lower = middle;
    }
}

static double evaluate_spline(int n, const double* x, const double* y, const double* y2, double t) {
    int lower, upper; double dx, a, b;
    spline_locate(n, x, t, &lower, &upper);
    dx = x[upper]-x[lower];
    a = (x[upper]-t)/dx;
    b = 1.0-a;
    return a*y[lower]+b*y[upper]+((a*a*a-a)*y2[lower]+(b*b*b-b)*y2[upper])*dx*dx/6.0;
}

static double evaluate_spline_derivative(int n, const double* x, const double* y, const double* y2, double t) {
    int lower, upper; double dx, a, b, dadx;
    spline_locate(n, x, t, &lower, &upper);
    dx = x[upper]-x[lower];
    a = (x[upper]-t)/dx;
    b = 1.0-a;
    dadx = -1.0/dx;
    return dadx*y[lower]-dadx*y[upper]+((1.0-3.0*a*a)*y2[lower]+(3.0*b*b-1.0)*y2[upper])*dx/6.0;
}

/* ------------------------------------------------------------------------------------------------------------- */
/* I4 lookup tables (AGBNPUtils.cpp)                                                                              */
/* ------------------------------------------------------------------------------------------------------------- */
/* AGBNPUtils.cpp:13-25 */
static double i4_switching_function(double x, double xa, double xb) {
    double d, u, u2, u3;
    if (x > xb) return 0.0;
    if (x < xa) return 1.0;
    d = 1.0/(xb - xa);
    u = (x - xa)*d;
    u2 = u*u;
    u3 = u*u2;
    return 1.0 - u3*(10.0-15.0*u+6.0*u2);
}

/* AGBNPUtils.cpp:27-32 */
static double i4_ogauss(double d2, double pi, double pj, double ai, double aj) {
    double deltai = 1.0/(ai+aj);
    double p = pi*pj;
    double kappa = exp(-ai*aj*d2*deltai);
    return p*kappa*pow(pi*deltai, 1.5);   /* sic: the reference passes the Gaussian prefactor `pi`, not M_PI */
}

/* AGBNPUtils.cpp:34-85 */
static double i4(double rij, double Ri, double Rj) {
    double u1, u2, u3, u4, u5, u6, a, u4sq, u5sq, q;
    double rij2 = rij*rij;
    const double twopi = 2.0*M_PI;
    const double twothirds = 2.0/3.0;
    if (rij > (Ri+Rj)) {
        u1 = rij+Rj; u2 = rij-Rj; u3 = u1*u2;
        u4 = 0.5*log(u1/u2);
        q = twopi*(Rj/u3 - u4/rij);
    } else {
        u1 = Rj-Ri;
        if (rij2 > u1*u1) {
            u1 = rij+Rj; u2 = rij-Rj; u3 = u1*u2;
            u4 = 1.0/u1; u4sq = u4*u4;
            u5 = 1.0/Ri; u5sq = u5*u5;
            u6 = 0.5*log(u1/Ri);
            q = twopi*(-(u4-u5) + (0.25*u3*(u4sq-u5sq) - u6)/rij);
        } else {
            if (Ri > Rj) {
                q = 0.0;
            } else {
                u1 = rij+Rj; u2 = Rj - rij; u3 = -u1*u2;
                if (rij < .001*Rj) {
                    a = rij/Rj;
                    u6 = (1.0 + twothirds*a*a)/Rj;
                    q = twopi*(2.0/Ri + Rj/u3 - u6);
                } else {
                    u6 = 0.5*log(u1/u2);
                    q = twopi*(2.0/Ri + Rj/u3 - u6/(rij));
                }
            }
        }
    }
    return q;
}

/* AGBNPUtils.cpp:87-97 */
static double i4ov(double rij, double Ri, double Rj, double gvol12_factor) {
    double ai = KFC/(Ri*Ri), pii = PFC, aj = KFC/(Rj*Rj), pjj = PFC;
    double d2 = rij*rij;
    double gvol = i4_ogauss(d2, pii, pjj, ai, aj);
    double volj = 4.0*M_PI*Rj*Rj*Rj/3.0;
    double newRj = pow((volj+gvol12_factor*gvol)/volj, 1.0/3.0)*Rj;
    return i4(rij, Ri, newRj);
}

typedef struct {
    int nti, ntj, nn;
    double *x, *y, *y2;          /* [nti*ntj][nn] */
    int *type_screened, *type_screener;
} i4_tables;

static int cmp_long(const void* a, const void* b) {
    long x = *(const long*)a, y = *(const long*)b;
    return x < y ? -1 : x > y;
}

/* unique radii under compare_pp10t (AGBNPUtils.h:173-179): keys long(r*10000); std::set keeps the FIRST inserted
 * representative of each class, iteration order is ascending key */
static int unique_radii(int n, const double* r, const int* use, long* keys, double* reps) {
    int i, k, m = 0;
    for (i = 0; i < n; i++) {
        long key;
        if (use && !use[i]) continue;
        key = (long)(r[i]*AGBNP_RADIUS_PRECISION);
        for (k = 0; k < m; k++) if (keys[k] == key) break;
        if (k == m) { keys[m] = key; reps[m] = r[i]; m++; }
    }
    /* sort classes by key, carrying representatives */
    {
        long* order = (long*) malloc(sizeof(long)*(size_t)(2*m+2));
        double* rr = (double*) malloc(sizeof(double)*(size_t)(m+1));
        for (k = 0; k < m; k++) { order[2*k] = keys[k]; order[2*k+1] = k; }
        qsort(order, (size_t)m, 2*sizeof(long), cmp_long);
        for (k = 0; k < m; k++) rr[k] = reps[order[2*k+1]];
        for (k = 0; k < m; k++) { keys[k] = order[2*k]; reps[k] = rr[k]; }
        free(order); free(rr);
    }
    return m;
}

static int find_type(int m, const long* keys, double r) {
    long key = (long)(r*AGBNP_RADIUS_PRECISION);
    int k;
    for (k = 0; k < m; k++) if (keys[k] == key) return k;
    return -1;
}

/* AGBNPUtils.cpp:134-200 (2D table) + :102-130 (one table) */
static i4_tables* i4_tables_create(int n, const double* radii, const int* ishydrogen) {
    i4_tables* T = (i4_tables*) calloc(1, sizeof(i4_tables));
    long *ki = (long*) malloc(sizeof(long)*(size_t)(n+1)), *kj = (long*) malloc(sizeof(long)*(size_t)(n+1));
    double *ri = (double*) malloc(sizeof(double)*(size_t)(n+1)), *rj = (double*) malloc(sizeof(double)*(size_t)(n+1));
    int* heavy = (int*) malloc(sizeof(int)*(size_t)(n+1));
    const int size = AGBNP_I4LOOKUP_NA;
    const double rmin = 0.0, rmax = AGBNP_I4LOOKUP_MAXA;
    int i, ti, tj, k;
    for (i = 0; i < n; i++) heavy[i] = !ishydrogen[i];
    T->nti = unique_radii(n, radii, NULL, ki, ri);
    T->ntj = unique_radii(n, radii, heavy, kj, rj);      /* roffset = 0.0 (AGBNPUtils.cpp:147) */
    T->nn = size;
    T->x = (double*) calloc((size_t)T->nti*T->ntj*size + 1, sizeof(double));
    T->y = (double*) calloc((size_t)T->nti*T->ntj*size + 1, sizeof(double));
    T->y2 = (double*) calloc((size_t)T->nti*T->ntj*size + 1, sizeof(double));
    for (ti = 0; ti < T->nti; ti++) for (tj = 0; tj < T->ntj; tj++) {
        double* x = T->x + ((size_t)ti*T->ntj+tj)*size;
        double* y = T->y + ((size_t)ti*T->ntj+tj)*size;
        double* y2 = T->y2 + ((size_t)ti*T->ntj+tj)*size;
        double dr = (rmax - rmin)/(size-1);
        double xa = 0.5*(rmax + rmin), xb = rmax;
        double gvol12_factor = 0.0;
        for (k = 0; k < size; k++) {
            double s;
            x[k] = k*dr + rmin;
            s = i4_switching_function(x[k], xa, xb);
            y[k] = s*i4ov(x[k], ri[ti], rj[tj], gvol12_factor);
        }
        create_natural_spline(size, x, y, y2);
    }
    T->type_screened = (int*) malloc(sizeof(int)*(size_t)(n+1));
    T->type_screener = (int*) malloc(sizeof(int)*(size_t)(n+1));
    for (i = 0; i < n; i++) {
        T->type_screened[i] = find_type(T->nti, ki, radii[i]);
        T->type_screener[i] = ishydrogen[i] ? -1 : find_type(T->ntj, kj, radii[i]);
    }
    free(ki); free(kj); free(ri); free(rj); free(heavy);
    return T;
}

static void i4_tables_destroy(i4_tables* T) {
    if (!T) return;
    free(T->x); free(T->y); free(T->y2); free(T->type_screened); free(T->type_screener); free(T);
}

static double i4_eval(const i4_tables* T, double d, int ti, int tj) {
    size_t o = ((size_t)ti*T->ntj+tj)*T->nn;
    return evaluate_spline(T->nn, T->x+o, T->y+o, T->y2+o, d);
}
static double i4_evalderiv(const i4_tables* T, double d, int ti, int tj) {
    size_t o = ((size_t)ti*T->ntj+tj)*T->nn;
    return evaluate_spline_derivative(T->nn, T->x+o, T->y+o, T->y2+o, d);
}

/* ------------------------------------------------------------------------------------------------------------- */
/* Reference kernel (ReferenceAGBNPKernels.cpp)                                                                   */
/* ------------------------------------------------------------------------------------------------------------- */
struct agbnp_oracle {
    int n, version, method;
    double cutoff, roffset, common_gamma;
    double *radii_vdw, *radii_large, *gammas, *vdw_alpha, *charge;
    int* ishydrogen;
    gtree tree;
    i4_tables* lut;
    /* outputs / by-products */
    double *free_volume, *self_volume, *free_volume_large, *self_volume_large, *vol_force, *vol_dv;
    double *vsf, *invbr, *invbr_fp, *br, *Y, *bru, *brw, *W, *U, *nu, *volumes;
    double scal[8];
    double counters[8];
};

/* float r2 < cutoff2 membership rule (shared bit-for-bit with the GPU) */
static int within_cutoff(const agbnp_oracle* h, const double* pos, int i, int j) {
    float dx, dy, dz, r2, c;
    if (h->method == 0) return 1;
    dx = (float)pos[3*j] - (float)pos[3*i];
    dy = (float)pos[3*j+1] - (float)pos[3*i+1];
    dz = (float)pos[3*j+2] - (float)pos[3*i+2];
    r2 = dx*dx + dy*dy + dz*dz;
    c = (float)h->cutoff;
    return r2 < c*c;
}

long agbnp_oracle_neighbor_pairs(int n, const float* pos, float cutoff, int* pairs, long max_pairs) {
    long cnt = 0;
    float c2 = cutoff*cutoff;
    int i, j;
    for (i = 0; i < n; i++) for (j = i+1; j < n; j++) {
        float dx = pos[3*j]-pos[3*i], dy = pos[3*j+1]-pos[3*i+1], dz = pos[3*j+2]-pos[3*i+2];
        float r2 = dx*dx + dy*dy + dz*dz;
        if (r2 < c2) {
            if (pairs && cnt < max_pairs) { pairs[2*cnt] = i; pairs[2*cnt+1] = j; }
            cnt++;
        }
    }
    return cnt;
}

/* ReferenceAGBNPKernels.cpp:41-55 */
static double agbnp_swf_invbr(double beta, double* fp) {
    const double a = 1.0/AGBNP_I4LOOKUP_MAXA;
    const double a2 = 1.0/(AGBNP_I4LOOKUP_MAXA*AGBNP_I4LOOKUP_MAXA);
    double t;
    if (beta < 0.0) { t = a; *fp = 0.0; }
    else { t = sqrt(a2 + beta*beta); *fp = beta/t; }
    return t;
}

static double* dalloc(int n) { return (double*) calloc((size_t)n+1, sizeof(double)); }

/* ReferenceAGBNPKernels.cpp:58-137 */
agbnp_oracle* agbnp_oracle_create(int version, int nonbonded_method, double cutoff, int n,
                                  const double* radius, const double* gamma, const double* alpha,
                                  const double* charge, const int* ishydrogen) {
    agbnp_oracle* h;
    int i;
    if (version < 0 || version > 2) { set_err("AGBNPForce::setVersion(): illegal version number"); return NULL; }
    if (version == 2) { set_err("oracle: AGBNP2 (version 2) is out of scope"); return NULL; }
    if (nonbonded_method == 2) { set_err("oracle: CutoffPeriodic is implemented nowhere in the reference"); return NULL; }
    if (nonbonded_method < 0 || nonbonded_method > 2) { set_err("oracle: bad nonbonded method"); return NULL; }
    h = (agbnp_oracle*) calloc(1, sizeof *h);
    h->n = n; h->version = version; h->method = nonbonded_method; h->cutoff = cutoff;
    h->roffset = AGBNP_RADIUS_INCREMENT;
    h->radii_vdw = dalloc(n); h->radii_large = dalloc(n); h->gammas = dalloc(n); h->vdw_alpha = dalloc(n);
    h->charge = dalloc(n); h->ishydrogen = (int*) calloc((size_t)n+1, sizeof(int));
    h->common_gamma = -1;
    for (i = 0; i < n; i++) {
        int hyd = ishydrogen[i] != 0;
        h->radii_large[i] = radius[i] + h->roffset;
        h->radii_vdw[i] = radius[i];
        h->gammas[i] = hyd ? 0.0 : gamma[i];
        h->vdw_alpha[i] = alpha[i];
        h->charge[i] = charge[i];
        h->ishydrogen[i] = hyd;
        if (h->common_gamma < 0 && !hyd) {
            h->common_gamma = gamma[i];
        } else if (!hyd && pow(h->common_gamma - gamma[i], 2) > FLT_MIN) {
            set_err("initialize(): AGBNP does not support multiple gamma values.");
            agbnp_oracle_destroy(h);
            return NULL;
        }
    }
    h->tree.natoms = n;
    h->lut = i4_tables_create(n, h->radii_vdw, h->ishydrogen);
    h->free_volume = dalloc(n); h->self_volume = dalloc(n); h->free_volume_large = dalloc(n);
    h->self_volume_large = dalloc(n); h->vol_force = dalloc(3*n); h->vol_dv = dalloc(n);
    h->vsf = dalloc(n); h->invbr = dalloc(n); h->invbr_fp = dalloc(n); h->br = dalloc(n); h->Y = dalloc(n);
    h->bru = dalloc(n); h->brw = dalloc(n); h->W = dalloc(n); h->U = dalloc(n); h->nu = dalloc(n);
    h->volumes = dalloc(n);
    return h;
}

void agbnp_oracle_destroy(agbnp_oracle* h) {
    if (!h) return;
    free(h->radii_vdw); free(h->radii_large); free(h->gammas); free(h->vdw_alpha); free(h->charge);
    free(h->ishydrogen); free(h->tree.ov); i4_tables_destroy(h->lut);
    free(h->free_volume); free(h->self_volume); free(h->free_volume_large); free(h->self_volume_large);
    free(h->vol_force); free(h->vol_dv); free(h->vsf); free(h->invbr); free(h->invbr_fp); free(h->br); free(h->Y);
    free(h->bru); free(h->brw); free(h->W); free(h->U); free(h->nu); free(h->volumes);
    free(h);
}

/* ReferenceAGBNPKernels.cpp:1796-1815 */
int agbnp_oracle_set_params(agbnp_oracle* h, const double* radius, const double* gamma, const double* alpha,
                            const double* charge, const int* ishydrogen) {
    int i;
    for (i = 0; i < h->n; i++) {
        int hyd = ishydrogen[i] != 0;
        if (pow(h->radii_vdw[i]-radius[i], 2) > 1.e-6) {
            set_err("updateParametersInContext: AGBNP plugin does not support changing atomic radii.");
            return -1;
        }
        if (hyd && h->ishydrogen[i] == 0) {
            set_err("updateParametersInContext: AGBNP plugin does not support changing heavy/hydrogen atoms.");
            return -1;
        }
        h->gammas[i] = hyd ? 0.0 : gamma[i];
        h->vdw_alpha[i] = alpha[i];
        h->charge[i] = charge[i];
    }
    return 0;
}

/* S1-S3: ReferenceAGBNPKernels.cpp:176-263 (v0) == :290-380 (v1) */
static double volume_terms(agbnp_oracle* h, const double* pos, double* force) {
    int i, n = h->n;
    double volume1, vol_energy1, volume2, vol_energy2;
    /* large radii */
    for (i = 0; i < n; i++) h->nu[i] = h->gammas[i]/h->roffset;
    for (i = 0; i < n; i++)
        h->volumes[i] = h->ishydrogen[i] > 0 ? 0.0 : 4.0*M_PI*pow(h->radii_large[i], 3)/3.0;
    compute_overlap_tree_r(&h->tree, pos, h->radii_large, h->volumes, h->nu, h->ishydrogen);
    compute_volume(&h->tree, h->volumes, &volume1, &vol_energy1, h->vol_force, h->vol_dv,
                   h->free_volume_large, h->self_volume_large);
    for (i = 0; i < 3*n; i++) force[i] += h->vol_force[i];
    h->counters[2] = (double) h->tree.n_eval2; h->counters[3] = (double) h->tree.n_eval3;
    h->counters[4] = (double) h->tree.size;
    /* vdW radii on the same topology */
    for (i = 0; i < n; i++) h->nu[i] = -h->gammas[i]/h->roffset;
    for (i = 0; i < n; i++)
        h->volumes[i] = h->ishydrogen[i] > 0 ? 0.0 : 4.0*M_PI*pow(h->radii_vdw[i], 3)/3.0;
    rescan_tree_v(&h->tree, pos, h->radii_vdw, h->volumes, h->nu, h->ishydrogen);
    compute_volume(&h->tree, h->volumes, &volume2, &vol_energy2, h->vol_force, h->vol_dv,
                   h->free_volume, h->self_volume);
    for (i = 0; i < 3*n; i++) force[i] += h->vol_force[i];
    h->scal[0] = vol_energy1; h->scal[1] = vol_energy2; h->scal[5] = volume1; h->scal[6] = volume2;
    return vol_energy1 + vol_energy2;
}

int agbnp_oracle_execute(agbnp_oracle* h, const double* pos, double* energy_out, double* force) {
    int i, j, k, n = h->n;
    double energy = 0.0;
    const double pifac = 1.0/(4.0*M_PI);
    const i4_tables* lut = h->lut;
    double tmpf[3];
    memset(h->scal, 0, sizeof h->scal);
    memset(h->counters, 0, sizeof h->counters);
    for (i = 0; i < 3*n; i++) force[i] = 0.0;

    energy += volume_terms(h, pos, force);
    if (h->version == 0) { *energy_out = energy; return 0; }     /* executeGVolSA :152-271 */

    /* S4 :421-430 */
    for (i = 0; i < n; i++) {
        double rad = h->radii_vdw[i];
        double vol = (4.0/3.0)*M_PI*rad*rad*rad;
        h->vsf[i] = h->self_volume[i]/vol;
    }
    /* S5 :437-454 */
    for (i = 0; i < n; i++) {
        double fp;
        h->invbr[i] = 1.0/h->radii_vdw[i];
        for (j = 0; j < n; j++) {
            double d;
            if (i == j) continue;
            if (h->ishydrogen[j] > 0) continue;
            if (!within_cutoff(h, pos, i, j)) continue;
            for (k = 0; k < 3; k++) tmpf[k] = pos[3*j+k] - pos[3*i+k];
            d = sqrt(tmpf[0]*tmpf[0] + tmpf[1]*tmpf[1] + tmpf[2]*tmpf[2]);
            if (d < AGBNP_I4LOOKUP_MAXA) {
                h->invbr[i] -= pifac*h->vsf[j]*i4_eval(lut, d, lut->type_screened[i], lut->type_screener[j]);
                h->counters[1] += 1;
            }
        }
        h->br[i] = 1.0/agbnp_swf_invbr(h->invbr[i], &fp);
        h->invbr_fp[i] = fp;
    }
    /* S6 :464-504 */
    {
        const double dielectric_in = 1.0, dielectric_out = 80.0;
        const double tokjmol = 4.184*332.0/10.0;
        const double dielectric_factor = tokjmol*(-0.5)*(1.0/dielectric_in - 1.0/dielectric_out);
        const double pt25 = 0.25;
        double gb_self_energy = 0.0, gb_pair_energy = 0.0, evdw = 0.0;
        for (i = 0; i < n; i++) h->Y[i] = 0.0;
        for (i = 0; i < n; i++) {
            double uself = dielectric_factor*h->charge[i]*h->charge[i]/h->br[i];
            gb_self_energy += uself;
            for (j = i+1; j < n; j++) {
                double dist[3], d2, qqf, qq, bb, etij, fgb, egb, fgb3, mw, ytij;
                if (!within_cutoff(h, pos, i, j)) continue;
                for (k = 0; k < 3; k++) dist[k] = pos[3*j+k] - pos[3*i+k];
                d2 = dist[0]*dist[0] + dist[1]*dist[1] + dist[2]*dist[2];
                qqf = h->charge[j]*h->charge[i];
                qq = dielectric_factor*qqf;
                bb = h->br[i]*h->br[j];
                etij = exp(-pt25*d2/bb);
                fgb = 1.0/sqrt(d2 + bb*etij);
                egb = 2.0*qq*fgb;
                gb_pair_energy += egb;
                fgb3 = fgb*fgb*fgb;
                mw = -2.0*qq*(1.0-pt25*etij)*fgb3;
                for (k = 0; k < 3; k++) { double g = dist[k]*mw; force[3*i+k] += g; force[3*j+k] -= g; }
                ytij = qqf*(bb+pt25*d2)*etij*fgb3;
                h->Y[i] += ytij;
                h->Y[j] += ytij;
                h->counters[0] += 1;
            }
        }
        energy += gb_pair_energy + gb_self_energy;
        h->scal[2] = gb_self_energy; h->scal[3] = gb_pair_energy;
        /* S7 :513-528 */
        for (i = 0; i < n; i++) evdw += h->vdw_alpha[i]/pow(h->br[i]+AGBNP_HB_RADIUS, 3);
        energy += evdw;
        h->scal[4] = evdw;
        for (i = 0; i < n; i++) {
            double br = h->br[i];
            h->brw[i] = -pifac*3.0*h->vdw_alpha[i]*br*br*h->invbr_fp[i]/pow(br+AGBNP_HB_RADIUS, 4);
        }
        /* S8 :537-542 */
        for (i = 0; i < n; i++) {
            double br = h->br[i], qi = h->charge[i];
            h->bru[i] = -pifac*dielectric_factor*(qi*qi + h->Y[i]*br)*h->invbr_fp[i];
        }
    }
    /* S9 :555-586 */
    for (i = 0; i < n; i++) h->W[i] = h->U[i] = 0.0;
    for (i = 0; i < n; i++) {
        for (j = 0; j < n; j++) {
            double dist[3], d, Qji = 0.0, dQji = 0.0;
            if (i == j) continue;
            if (h->ishydrogen[j] > 0) continue;
            if (!within_cutoff(h, pos, i, j)) continue;
            for (k = 0; k < 3; k++) dist[k] = pos[3*j+k] - pos[3*i+k];
            d = sqrt(dist[0]*dist[0] + dist[1]*dist[1] + dist[2]*dist[2]);
            if (d < AGBNP_I4LOOKUP_MAXA) {
                int ti = lut->type_screened[i], tj = lut->type_screener[j];
                Qji = i4_eval(lut, d, ti, tj);
                dQji = i4_evalderiv(lut, d, ti, tj);
            }
            /* `Vec3 / double` multiplies by the reciprocal (OpenMM Vec3.h, mirrored in oracle/shim) */
            h->W[j] += h->brw[i]*Qji;
            for (k = 0; k < 3; k++) {
                double w = dist[k]*h->brw[i]*h->vsf[j]*dQji*(1.0/d);
                force[3*i+k] += w; force[3*j+k] -= w;
            }
            h->U[j] += h->bru[i]*Qji;
            for (k = 0; k < 3; k++) {
                double w = dist[k]*h->bru[i]*h->vsf[j]*dQji*(1.0/d);
                force[3*i+k] += w; force[3*j+k] -= w;
            }
        }
    }
    /* S10 :718-727, S11 :738-747 */
    {
        double volume_tmp, vol_energy_tmp;
        int pass;
        for (pass = 0; pass < 2; pass++) {
            const double* src = pass == 0 ? h->W : h->U;
            for (i = 0; i < n; i++) {
                double vol = 4.0*M_PI*pow(h->radii_vdw[i], 3)/3.0;
                h->nu[i] = src[i]/vol;
            }
            rescan_tree_g(&h->tree, h->nu);
            compute_volume(&h->tree, h->volumes, &volume_tmp, &vol_energy_tmp, h->vol_force, h->vol_dv,
                           h->free_volume, h->self_volume);
            for (i = 0; i < 3*n; i++) force[i] += h->vol_force[i];
        }
    }
    *energy_out = energy;
    return 0;
}

int agbnp_oracle_get(agbnp_oracle* h, int what, double* out) {
    const double* src = NULL;
    int i;
    switch (what) {
    case 0: src = h->self_volume; break;
    case 1: src = h->self_volume_large; break;
    case 2: src = h->vsf; break;
    case 3: src = h->br; break;
    case 4: src = h->invbr_fp; break;
    case 5: src = h->Y; break;
    case 6: src = h->bru; break;
    case 7: src = h->brw; break;
    case 8: src = h->W; break;
    case 9: src = h->U; break;
    case 10: for (i = 0; i < h->n; i++) out[i] = h->lut->type_screened[i]; return 0;
    case 11: for (i = 0; i < h->n; i++) out[i] = h->lut->type_screener[i]; return 0;
    case 12: src = h->free_volume; break;
    case 13: src = h->free_volume_large; break;
    default: set_err("agbnp_oracle_get: bad selector"); return -1;
    }
    memcpy(out, src, sizeof(double)*(size_t)h->n);
    return 0;
}

double agbnp_oracle_scalar(agbnp_oracle* h, int what) { return (what >= 0 && what < 8) ? h->scal[what] : 0.0; }
double agbnp_oracle_counter(agbnp_oracle* h, int what) { return (what >= 0 && what < 8) ? h->counters[what] : 0.0; }

int agbnp_oracle_tree_size(agbnp_oracle* h) { return h->tree.size; }

int agbnp_oracle_tree_dump(agbnp_oracle* h, int* level, int* atom, int* parent, int* child_start, int* child_count,
                           double* volume, double* gvol) {
    int s;
    for (s = 0; s < h->tree.size; s++) {
        const goverlap* o = &h->tree.ov[s];
        if (level) level[s] = o->level;
        if (atom) atom[s] = o->atom;
        if (parent) parent[s] = o->parent_index;
        if (child_start) child_start[s] = o->children_startindex;
        if (child_count) child_count[s] = o->children_count;
        if (volume) volume[s] = o->volume;
        if (gvol) gvol[s] = o->g.v;
    }
    return 0;
}

int agbnp_oracle_i4_dims(agbnp_oracle* h, int* nti, int* ntj, int* nn) {
    *nti = h->lut->nti; *ntj = h->lut->ntj; *nn = h->lut->nn;
    return 0;
}

int agbnp_oracle_i4_table(agbnp_oracle* h, int ti, int tj, double* x, double* y, double* y2) {
    size_t o = ((size_t)ti*h->lut->ntj+tj)*h->lut->nn;
    memcpy(x, h->lut->x+o, sizeof(double)*(size_t)h->lut->nn);
    memcpy(y, h->lut->y+o, sizeof(double)*(size_t)h->lut->nn);
    memcpy(y2, h->lut->y2+o, sizeof(double)*(size_t)h->lut->nn);
    return 0;
}
