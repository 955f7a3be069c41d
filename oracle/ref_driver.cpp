// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI driver around the reference's own, UNMODIFIED sources (compiled where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libagbnp_ref.so):
//     gaussvol/gaussvol.cpp, openmmapi/src/AGBNPForce.cpp, openmmapi/src/AGBNPUtils.cpp,
//     platforms/reference/src/ReferenceAGBNPKernels.cpp
// against the small OpenMM shim in oracle/shim/.  It exposes
//   * the Reference-platform kernel (ReferenceCalcAGBNPForceKernel::initialize/execute/copyParametersToContext,
//     platforms/reference/src/ReferenceAGBNPKernels.cpp:58-149,1796-1815) and its per-atom by-products, and
//   * the GaussVol overlap tree (gaussvol/gaussvol.h:123-312) for topology / self-volume parity checks.
// `#define private public` below only widens access for reading diagnostics; class layout is unchanged.

#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <set>
#include <iostream>
#include <sstream>
#include <algorithm>

#define private public
#include "AGBNPForce.h"
#include "AGBNPKernels.h"
#include "AGBNPUtils.h"
#include "ReferenceAGBNPKernels.h"
#include "internal/AGBNPForceImpl.h"
#include "gaussvol.h"
#undef private

using namespace OpenMM;
using namespace AGBNPPlugin;
using std::vector;

// ---- members the reference declares but whose translation units (OpenMM proper, AGBNPForceImpl.cpp) are not built ----
namespace OpenMM {
ForceImpl& Force::getImplInContext(Context&) { throw OpenMMException("oracle shim: no Context"); }
ContextImpl& Force::getContextImpl(Context&) { throw OpenMMException("oracle shim: no Context"); }
}
namespace AGBNPPlugin {
AGBNPForceImpl::AGBNPForceImpl(const AGBNPForce& owner) : owner(owner) {}
AGBNPForceImpl::~AGBNPForceImpl() {}
void AGBNPForceImpl::initialize(ContextImpl&) {}
double AGBNPForceImpl::calcForcesAndEnergy(ContextImpl&, bool, bool, int) { return 0.0; }
std::vector<std::string> AGBNPForceImpl::getKernelNames() { return std::vector<std::string>(); }
void AGBNPForceImpl::updateParametersInContext(ContextImpl&) {}
}

namespace {
std::string g_err;

struct RefHandle {
    AGBNPForce force;
    Platform platform;
    ReferenceCalcAGBNPForceKernel* kernel;
    ContextImpl ctx;
    ReferencePlatform::PlatformData pd;
    vector<RealVec> pos, frc;
    int n;
    RefHandle() : kernel(0), n(0) {}
    ~RefHandle() { delete kernel; }
};

struct TreeHandle {
    GaussVol* gv;
    vector<int> ish;
    int n;
    TreeHandle() : gv(0), n(0) {}
    ~TreeHandle() { delete gv; }
};
}

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// AGBNPForce surface: setVersion range check (openmmapi/src/AGBNPForce.cpp:52-59).  Returns 0 / -1.
int ref_check_version(int version) {
    try { AGBNPForce f; f.setVersion(version); return 0; }
    catch (const std::exception& e) { g_err = e.what(); return -1; }
}

void* ref_create(int version, int n, const double* radius, const double* gamma, const double* alpha,
                 const double* charge, const int* ishydrogen) {
    RefHandle* h = new RefHandle();
    try {
        h->n = n;
        h->force.setVersion(version);
        for (int i = 0; i < n; i++)
            h->force.addParticle(radius[i], gamma[i], alpha[i], charge[i], ishydrogen[i] != 0);
        h->kernel = new ReferenceCalcAGBNPForceKernel(CalcAGBNPForceKernel::Name(), h->platform);
        h->kernel->initialize(OpenMM::System(), h->force);
        h->pos.resize(n); h->frc.resize(n);
        h->pd.positions = &h->pos; h->pd.forces = &h->frc;
        h->ctx.setPlatformData(&h->pd);
    } catch (const std::exception& e) { g_err = e.what(); delete h; return 0; }
    return h;
}

void ref_destroy(void* hv) { delete (RefHandle*) hv; }

// copyParametersToContext (ReferenceAGBNPKernels.cpp:1796-1815)
int ref_set_params(void* hv, const double* radius, const double* gamma, const double* alpha, const double* charge,
                   const int* ishydrogen) {
    RefHandle* h = (RefHandle*) hv;
    try {
        for (int i = 0; i < h->n; i++)
            h->force.setParticleParameters(i, radius[i], gamma[i], alpha[i], charge[i], ishydrogen[i] != 0);
        h->kernel->copyParametersToContext(h->ctx, h->force);
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

// execute (ReferenceAGBNPKernels.cpp:139-149): energy returned, forces accumulated into a zeroed buffer.
int ref_execute(void* hv, const double* pos, double* energy, double* forces) {
    RefHandle* h = (RefHandle*) hv;
    try {
        for (int i = 0; i < h->n; i++) {
            h->pos[i] = RealVec(pos[3*i], pos[3*i+1], pos[3*i+2]);
            h->frc[i] = RealVec(0, 0, 0);
        }
        // executeAGBNP2 prints; v0/v1 are silent (verbose_level = 0)
        double e = h->kernel->execute(h->ctx, true, true);
        if (energy) *energy = e;
        if (forces) for (int i = 0; i < h->n; i++) for (int k = 0; k < 3; k++) forces[3*i+k] = h->frc[i][k];
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

// per-atom by-products after the last ref_execute: what = 0 self_volume (vdW radii), 1 volume_scaling_factor,
// 2 born_radius, 3 inverse_born_radius_fp, 4 radius_type_screened, 5 radius_type_screener
int ref_get(void* hv, int what, double* out) {
    RefHandle* h = (RefHandle*) hv;
    ReferenceCalcAGBNPForceKernel* k = h->kernel;
    for (int i = 0; i < h->n; i++) {
        switch (what) {
        case 0: out[i] = k->self_volume[i]; break;
        case 1: out[i] = k->volume_scaling_factor[i]; break;
        case 2: out[i] = k->born_radius[i]; break;
        case 3: out[i] = k->inverse_born_radius_fp[i]; break;
        case 4: out[i] = k->i4_lut->radius_type_screened[i]; break;
        case 5: out[i] = k->i4_lut->radius_type_screener[i]; break;
        default: g_err = "ref_get: bad selector"; return -1;
        }
    }
    return 0;
}

// I4 lookup tables as the Reference kernel built them (AGBNPUtils.cpp:134-200): returns ntypes via pointers and
// copies node arrays (x,y,y2) of table (ti,tj) when out pointers are non-null.
int ref_i4_dims(void* hv, int* ntypes_screened, int* ntypes_screener, int* nnodes) {
    RefHandle* h = (RefHandle*) hv;
    *ntypes_screened = h->kernel->i4_lut->ntypes_screened;
    *ntypes_screener = h->kernel->i4_lut->ntypes_screener;
    *nnodes = (int) h->kernel->i4_lut->tables[0]->table->xt.size();
    return 0;
}
int ref_i4_table(void* hv, int ti, int tj, double* x, double* y, double* y2) {
    RefHandle* h = (RefHandle*) hv;
    AGBNPI42DLookupTable* t = h->kernel->i4_lut;
    AGBNPLookupTable* tb = t->tables[ti*t->ntypes_screener+tj]->table;
    for (size_t k = 0; k < tb->xt.size(); k++) { x[k] = tb->xt[k]; y[k] = tb->yt[k]; y2[k] = tb->y2t[k]; }
    return 0;
}
int ref_i4_eval(void* hv, double d, int ti, int tj, double* q, double* dq) {
    RefHandle* h = (RefHandle*) hv;
    try {
        *q = h->kernel->i4_lut->eval(d, ti, tj);
        *dq = h->kernel->i4_lut->evalderiv(d, ti, tj);
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

// overlap tree left inside the kernel after execute (vdW-radius volumes, topology from the large radii)
int ref_kernel_tree_size(void* hv) {
    RefHandle* h = (RefHandle*) hv;
    return (int) h->kernel->gvol->tree->overlaps.size();
}

// ---- direct GaussVol access (gaussvol/gaussvol.h:205-312) ----
void* gv_create(int n, const int* ishydrogen) {
    TreeHandle* t = new TreeHandle();
    t->n = n;
    t->ish.assign(ishydrogen, ishydrogen+n);
    t->gv = new GaussVol(n, t->ish);
    return t;
}
void gv_destroy(void* tv) { delete (TreeHandle*) tv; }

static void to_vec(int n, const double* p, vector<RealVec>& v) {
    v.resize(n);
    for (int i = 0; i < n; i++) v[i] = RealVec(p[3*i], p[3*i+1], p[3*i+2]);
}

// mode 0: compute_tree (build); 1: rescan_tree_volumes; 2: rescan_tree_gammas.  Radii/volumes/gammas are set first.
int gv_update(void* tv, int mode, const double* pos, const double* radii, const double* volumes, const double* gammas) {
    TreeHandle* t = (TreeHandle*) tv;
    try {
        vector<RealVec> p; to_vec(t->n, pos, p);
        vector<RealOpenMM> r(radii, radii+t->n), v(volumes, volumes+t->n), g(gammas, gammas+t->n);
        t->gv->setRadii(r); t->gv->setVolumes(v); t->gv->setGammas(g);
        if (mode == 0) t->gv->compute_tree(p);
        else if (mode == 1) t->gv->rescan_tree_volumes(p);
        else t->gv->rescan_tree_gammas();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

// compute_volume (gaussvol.cpp:589-606)
int gv_volume(void* tv, const double* pos, double* volume, double* energy, double* force, double* gradV,
              double* free_volume, double* self_volume) {
    TreeHandle* t = (TreeHandle*) tv;
    try {
        int n = t->n;
        vector<RealVec> p; to_vec(n, pos, p);
        vector<RealVec> f(n);
        vector<RealOpenMM> gv(n), fv(n), sv(n);
        RealOpenMM vol, en;
        t->gv->compute_volume(p, vol, en, f, gv, fv, sv);
        if (volume) *volume = vol;
        if (energy) *energy = en;
        for (int i = 0; i < n; i++) {
            if (force) for (int k = 0; k < 3; k++) force[3*i+k] = f[i][k];
            if (gradV) gradV[i] = gv[i];
            if (free_volume) free_volume[i] = fv[i];
            if (self_volume) self_volume[i] = sv[i];
        }
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

int gv_tree_size(void* tv) { return (int) ((TreeHandle*) tv)->gv->tree->overlaps.size(); }

// flat dump of the tree (slot order = the reference's DFS-append order)
int gv_tree_dump(void* tv, int* level, int* atom, int* parent, int* child_start, int* child_count,
                 double* volume, double* gvol, double* gamma1i, double* ga, double* gc /*3 per slot*/,
                 double* dv1 /*3 per slot*/, double* dvv1, double* sfp) {
    vector<GOverlap>& ov = ((TreeHandle*) tv)->gv->tree->overlaps;
    for (size_t s = 0; s < ov.size(); s++) {
        if (level) level[s] = ov[s].level;
        if (atom) atom[s] = ov[s].atom;
        if (parent) parent[s] = ov[s].parent_index;
        if (child_start) child_start[s] = ov[s].children_startindex;
        if (child_count) child_count[s] = ov[s].children_count;
        if (volume) volume[s] = ov[s].volume;
        if (gvol) gvol[s] = ov[s].g.v;
        if (gamma1i) gamma1i[s] = ov[s].gamma1i;
        if (ga) ga[s] = ov[s].g.a;
        if (gc) for (int k = 0; k < 3; k++) gc[3*s+k] = ov[s].g.c[k];
        if (dv1) for (int k = 0; k < 3; k++) dv1[3*s+k] = ov[s].dv1[k];
        if (dvv1) dvv1[s] = ov[s].dvv1;
        if (sfp) sfp[s] = ov[s].sfp;
    }
    return 0;
}

// float-literal constants exactly as the reference's translation units see them (gaussvol.h:46-63, AGBNPForce.h:25-33)
void ref_constants(double* out /*8*/) {
    out[0] = KFC; out[1] = VOLMINA; out[2] = VOLMINB; out[3] = MIN_GVOL;
    out[4] = AGBNP_RADIUS_INCREMENT; out[5] = AGBNP_HB_RADIUS; out[6] = AGBNP_I4LOOKUP_MAXA; out[7] = MAX_ORDER;
}

} // extern "C"
