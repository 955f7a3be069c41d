// Host-side setup for the B200 AGBNP1/GaussVol path (see agbnp_setup.h).
#include "agbnp_setup.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <map>
#include <numeric>

namespace agbnp_b200_impl {

Constants Constants::make() {
    Constants c;
    const float ang = 0.1f, ang3 = 0.001f;
    c.kfc = (double) 2.2269859253f;                 // gaussvol.h:46
    c.volmina = (double) (0.01f*ang3);              // gaussvol.h:62 (float product, then promoted)
    c.volminb = (double) (0.1f*ang3);               // gaussvol.h:63
    c.min_gvol = (double) FLT_MIN;                  // gaussvol.h:52
    c.roffset = (double) (0.5f*ang);                // AGBNPForce.h:25
    c.hb_radius = 1.4*(double) ang;                 // AGBNPForce.h:33
    c.i4_maxa = 2.0;                                // AGBNPUtils.h:124
    c.i4_nodes = 16;                                // AGBNPUtils.h:126
    c.max_order = 8;                                // gaussvol.h:55
    const double tokjmol = 4.184*332.0/10.0;        // ReferenceAGBNPKernels.cpp:467
    c.dielectric_factor = tokjmol*(-0.5)*(1.0/1.0 - 1.0/80.0);   // :468
    return c;
}

// ---------------------------------------------------------------------------------------------------------------
// Q4 pair descreening integral and its tabulation (reference: openmmapi/src/AGBNPUtils.cpp:13-130)
// ---------------------------------------------------------------------------------------------------------------
namespace {

// quintic switch, 1 below xa, 0 above xb (AGBNPUtils.cpp:13-25)
double q4_switch(double x, double xa, double xb) {
    if (x > xb) return 0.0;
    if (x < xa) return 1.0;
    const double u = (x-xa)/(xb-xa);
    const double u3 = u*u*u;
    return 1.0 - u3*(10.0 - 15.0*u + 6.0*u*u);
}

// closed-form integral of 1/r^4 over the sphere of radius Rj centred at distance rij, outside the sphere Ri of the
// screened atom; three geometric regimes (AGBNPUtils.cpp:34-85)
double q4_integral(double rij, double Ri, double Rj) {
    const double twopi = 2.0*M_PI;
    if (rij > Ri+Rj) {                                  // separated spheres
        const double up = rij+Rj, um = rij-Rj;
        return twopi*(Rj/(up*um) - 0.5*std::log(up/um)/rij);
    }
    const double dR = Rj-Ri;
    if (rij*rij > dR*dR) {                              // partial overlap
        const double up = rij+Rj, um = rij-Rj;
        const double iu = 1.0/up, iR = 1.0/Ri;
        return twopi*(-(iu-iR) + (0.25*up*um*(iu*iu-iR*iR) - 0.5*std::log(up/Ri))/rij);
    }
    if (Ri > Rj) return 0.0;                            // screener entirely inside the screened atom
    const double up = rij+Rj, um = Rj-rij;              // screened atom inside the screener
    const double u3 = -up*um;
    if (rij < 0.001*Rj) {                               // series at rij -> 0
        const double a = rij/Rj;
        return twopi*(2.0/Ri + Rj/u3 - (1.0 + (2.0/3.0)*a*a)/Rj);
    }
    return twopi*(2.0/Ri + Rj/u3 - 0.5*std::log(up/um)/rij);
}

// natural cubic spline second derivatives on arbitrary nodes (OpenMM SplineFitter::createNaturalSpline as called at
// AGBNPUtils.h:104): tridiagonal solve with zero curvature at both ends
void natural_spline(const std::vector<double>& x, const std::vector<double>& y, std::vector<double>& y2) {
    const int n = (int) x.size();
    y2.assign(n, 0.0);
    if (n <= 2) return;
    std::vector<double> sub(n, 0.0), diag(n, 1.0), sup(n, 0.0), rhs(n, 0.0), g(n, 0.0);
    for (int i = 1; i < n-1; i++) {
        sub[i] = x[i]-x[i-1];
        diag[i] = 2.0*(x[i+1]-x[i-1]);
        sup[i] = x[i+1]-x[i];
        rhs[i] = 6.0*((y[i+1]-y[i])/(x[i+1]-x[i]) - (y[i]-y[i-1])/(x[i]-x[i-1]));
    }
    y2[0] = rhs[0]/diag[0];
    double beta = diag[0];
    for (int i = 1; i < n; i++) {
        g[i] = sup[i-1]/beta;
        beta = diag[i]-sub[i]*g[i];
        y2[i] = (rhs[i]-sub[i]*y2[i-1])/beta;
    }
    for (int i = n-2; i >= 0; i--) y2[i] -= g[i+1]*y2[i+1];
}

// radius classes at 1e-4 nm resolution, truncating (AGBNPUtils.h:155,173-179); the class representative is the first
// radius seen (std::set::insert semantics), classes ordered by key
struct RadiusClasses {
    std::vector<long> key;
    std::vector<double> rep;
    void build(const std::vector<double>& r, const std::vector<int>* mask) {
        std::map<long, double> m;
        for (size_t i = 0; i < r.size(); i++) {
            if (mask && !(*mask)[i]) continue;
            const long k = (long) (r[i]*10000);
            if (!m.count(k)) m[k] = r[i];
        }
        for (auto& kv : m) { key.push_back(kv.first); rep.push_back(kv.second); }
    }
    int find(double r) const {
        const long k = (long) (r*10000);
        auto it = std::lower_bound(key.begin(), key.end(), k);
        return (it != key.end() && *it == k) ? (int) (it-key.begin()) : -1;
    }
};

} // namespace

void I4Tables::build(const std::vector<double>& radii, const std::vector<int>& ishydrogen, const Constants& c) {
    const int n = (int) radii.size();
    std::vector<int> heavy(n);
    for (int i = 0; i < n; i++) heavy[i] = !ishydrogen[i];
    RadiusClasses ci, cj;
    ci.build(radii, nullptr);        // screened: every atom's vdW radius (AGBNPUtils.cpp:142-144)
    cj.build(radii, &heavy);         // screeners: heavy atoms only, no radius offset for AGBNP1 (:147-151)
    ntypes_screened = (int) ci.rep.size();
    ntypes_screener = (int) cj.rep.size();
    nodes = c.i4_nodes;
    const double rmin = 0.0, rmax = c.i4_maxa;
    h = (rmax-rmin)/(nodes-1);
    const double xa = 0.5*(rmax+rmin), xb = rmax;       // switch from the midpoint to the end (:116-117)
    y.assign((size_t) ntypes_screened*ntypes_screener*nodes, 0.0);
    y2 = y;
    packed.assign((size_t) ntypes_screened*ntypes_screener*(nodes-1)*4, 0.f);
    std::vector<double> x(nodes), yy(nodes), d2;
    for (int ti = 0; ti < ntypes_screened; ti++) for (int tj = 0; tj < ntypes_screener; tj++) {
        for (int k = 0; k < nodes; k++) {
            x[k] = k*h + rmin;
            // gvol12_factor = 0 for AGBNP1 (:121), so the "overlap-corrected" radius of i4ov (:87-97) is Rj itself
            yy[k] = q4_switch(x[k], xa, xb)*q4_integral(x[k], ci.rep[ti], cj.rep[tj]);
        }
        natural_spline(x, yy, d2);
        const size_t o = ((size_t) ti*ntypes_screener+tj)*nodes;
        for (int k = 0; k < nodes; k++) { y[o+k] = yy[k]; y2[o+k] = d2[k]; }
        float* p = &packed[((size_t) ti*ntypes_screener+tj)*(nodes-1)*4];
        for (int k = 0; k < nodes-1; k++) {
            p[4*k+0] = (float) yy[k];
            p[4*k+1] = (float) yy[k+1];
            p[4*k+2] = (float) (d2[k]*h*h/6.0);
            p[4*k+3] = (float) (d2[k+1]*h*h/6.0);
        }
    }
    type_screened.resize(n);
    type_screener.resize(n);
    for (int i = 0; i < n; i++) {
        type_screened[i] = ci.find(radii[i]);
        type_screener[i] = ishydrogen[i] ? -1 : cj.find(radii[i]);
    }
}

double I4Tables::eval(double d, int ti, int tj) const {
    int k = std::min((int) (d/h), nodes-2);
    const size_t o = ((size_t) ti*ntypes_screener+tj)*nodes;
    const double a = ((k+1)*h - d)/h, b = 1.0-a;
    return a*y[o+k] + b*y[o+k+1] + ((a*a*a-a)*y2[o+k] + (b*b*b-b)*y2[o+k+1])*h*h/6.0;
}

double I4Tables::evalderiv(double d, int ti, int tj) const {
    int k = std::min((int) (d/h), nodes-2);
    const size_t o = ((size_t) ti*ntypes_screener+tj)*nodes;
    const double a = ((k+1)*h - d)/h, b = 1.0-a;
    return (y[o+k+1]-y[o+k])/h + ((1.0-3.0*a*a)*y2[o+k] + (3.0*b*b-1.0)*y2[o+k+1])*h/6.0;
}

// ---------------------------------------------------------------------------------------------------------------
// parameter unpacking (ReferenceAGBNPKernels.cpp:58-137) and update (:1796-1815)
// ---------------------------------------------------------------------------------------------------------------
std::string SystemParams::init(int version_, int n_, const double* radius_, const double* gamma_, const double* alpha_,
                               const double* charge_, const unsigned char* ish_, const Constants& c) {
    n = n_;
    version = version_;
    radius.assign(radius_, radius_+n);
    alpha.assign(alpha_, alpha_+n);
    charge.assign(charge_, charge_+n);
    gamma.resize(n);
    ishydrogen.resize(n);
    common_gamma = -1;
    for (int i = 0; i < n; i++) {
        const bool hyd = ish_[i] != 0;
        ishydrogen[i] = hyd ? 1 : 0;
        gamma[i] = hyd ? 0.0 : gamma_[i];
        if (common_gamma < 0 && !hyd) common_gamma = gamma_[i];
        else if (!hyd && std::pow(common_gamma-gamma_[i], 2) > FLT_MIN)
            return "initialize(): AGBNP does not support multiple gamma values.";
    }
    aL.resize(n); vL.resize(n); aS.resize(n); vS.resize(n);
    double rl_min = 1e30, rl_max = 0;
    for (int i = 0; i < n; i++) {
        const double rs = radius[i], rl = radius[i]+c.roffset;
        aL[i] = c.kfc/(rl*rl);
        aS[i] = c.kfc/(rs*rs);
        vL[i] = ishydrogen[i] ? 0.0 : 4.0*M_PI*std::pow(rl, 3)/3.0;
        vS[i] = ishydrogen[i] ? 0.0 : 4.0*M_PI*std::pow(rs, 3)/3.0;
        if (!ishydrogen[i]) { rl_min = std::min(rl_min, rl); rl_max = std::max(rl_max, rl); }
    }
    // level-2 pair-list filter on upward-binned radii
    const int maxbins = 64;
    if (rl_max < rl_min) { rl_min = rl_max = 0.1; }
    double w = std::max(RC_BIN_WIDTH, (rl_max-rl_min)/(maxbins-1));
    nbins = std::min(maxbins, (int) std::ceil((rl_max-rl_min)/w - 1e-9) + 1);
    rc_bin.assign(n, 0);
    for (int i = 0; i < n; i++) {
        if (ishydrogen[i]) continue;
        int b = (int) std::ceil((radius[i]+c.roffset-rl_min)/w - 1e-9);
        rc_bin[i] = std::min(std::max(b, 0), nbins-1);
    }
    rc2.assign((size_t) nbins*nbins, 0.f);
    rc2max.assign(nbins, 0.f);
    for (int bi = 0; bi < nbins; bi++) for (int bj = 0; bj < nbins; bj++) {
        const double ri = rl_min+bi*w+1e-9, rj = rl_min+bj*w+1e-9;      // upper edges: overlaps grow with radius
        const double ai = c.kfc/(ri*ri), aj = c.kfc/(rj*rj);
        const double vi = 4.0*M_PI*ri*ri*ri/3.0, vj = 4.0*M_PI*rj*rj*rj/3.0;
        const double df = ai*aj/(ai+aj);
        const double pref = vi*vj*std::pow(df/M_PI, 1.5);
        double lim = pref > c.volmina ? std::log(pref/c.volmina)/df : 0.0;
        lim = 1.02*lim + 1e-6;                                        // margin over float rounding of r2
        rc2[(size_t) bi*nbins+bj] = (float) lim;
        rc2max[bi] = std::max(rc2max[bi], (float) lim);
    }
    i4.build(radius, ishydrogen, c);
    return "";
}

std::string SystemParams::update(int n_, const double* radius_, const double* gamma_, const double* alpha_,
                                 const double* charge_, const unsigned char* ish_) {
    if (n_ != n) return "updateParametersInContext: The number of AGBNP particles has changed";
    for (int i = 0; i < n; i++) {
        const bool hyd = ish_[i] != 0;
        if (std::pow(radius[i]-radius_[i], 2) > 1.e-6)
            return "updateParametersInContext: AGBNP plugin does not support changing atomic radii.";
        if (hyd && ishydrogen[i] == 0)
            return "updateParametersInContext: AGBNP plugin does not support changing heavy/hydrogen atoms.";
    }
    for (int i = 0; i < n; i++) {
        const bool hyd = ish_[i] != 0;
        gamma[i] = hyd ? 0.0 : gamma_[i];
        alpha[i] = alpha_[i];
        charge[i] = charge_[i];
    }
    return "";
}

// ---------------------------------------------------------------------------------------------------------------
// spatial ordering
// ---------------------------------------------------------------------------------------------------------------
namespace {
inline uint64_t spread3(uint32_t v) {           // 21 bits -> every third bit
    uint64_t x = v & 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffULL;
    x = (x | x << 16) & 0x1f0000ff0000ffULL;
    x = (x | x << 8) & 0x100f00f00f00f00fULL;
    x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
    x = (x | x << 2) & 0x1249249249249249ULL;
    return x;
}
}

void morton_order(const float* xyz, int stride, const std::vector<int>& ishydrogen,
                  std::vector<int>& heavy_sorted, std::vector<int>& hydrogen_sorted) {
    const int n = (int) ishydrogen.size();
    float lo[3] = {1e30f, 1e30f, 1e30f};
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) lo[k] = std::min(lo[k], xyz[(size_t) i*stride+k]);
    const float cell = 0.4f;   // nm; blocks of 32 consecutive atoms then span roughly (0.8 nm)^3
    std::vector<std::pair<uint64_t, int>> hv, hy;
    for (int i = 0; i < n; i++) {
        uint32_t c[3];
        for (int k = 0; k < 3; k++) {
            float f = (xyz[(size_t) i*stride+k]-lo[k])/cell;
            c[k] = (uint32_t) std::min(std::max(f, 0.f), 2097151.f);
        }
        const uint64_t code = spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2);
        (ishydrogen[i] ? hy : hv).push_back({code, i});
    }
    std::sort(hv.begin(), hv.end());
    std::sort(hy.begin(), hy.end());
    heavy_sorted.clear(); hydrogen_sorted.clear();
    for (auto& p : hv) heavy_sorted.push_back(p.second);
    for (auto& p : hy) hydrogen_sorted.push_back(p.second);
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <emmintrin.h>
#define AGBNP_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define AGBNP_CLONES
#endif

void pack_positions(const double* pos, float* posq, int n) {
#if defined(__x86_64__) && defined(__GNUC__)
    for (int i = 0; i < n; i++) {
        const __m128 xy = _mm_cvtpd_ps(_mm_loadu_pd(pos + 3*(size_t) i));         // x, y, 0, 0
        const __m128 z = _mm_cvtpd_ps(_mm_load_sd(pos + 3*(size_t) i + 2));         // z, 0, 0, 0
        _mm_storeu_ps(posq + 4*(size_t) i, _mm_movelh_ps(xy, z));                  // x, y, z, 0
    }
#else
    for (int i = 0; i < n; i++) {
        posq[4*i] = (float) pos[3*i]; posq[4*i+1] = (float) pos[3*i+1]; posq[4*i+2] = (float) pos[3*i+2]; posq[4*i+3] = 0.f;
    }
#endif
}

AGBNP_CLONES void add_forces(const float* __restrict__ src, double* __restrict__ dst, int n3) {
    for (int i = 0; i < n3; i++) dst[i] += (double) src[i];
}

AGBNP_CLONES void set_forces(const float* __restrict__ src, double* __restrict__ dst, int n3) {
    for (int i = 0; i < n3; i++) dst[i] = (double) src[i];
}

} // namespace agbnp_b200_impl
