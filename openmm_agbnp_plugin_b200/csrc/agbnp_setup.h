// Host-side setup for the B200 AGBNP1/GaussVol path: everything CalcAGBNPForceKernel::initialize does before the first
// execute (reference: platforms/reference/src/ReferenceAGBNPKernels.cpp:58-137, openmmapi/src/AGBNPUtils.cpp:13-214).
// Pure C++ (no CUDA); the device driver in agbnp_b200.cu uploads what is built here.
#ifndef AGBNP_SETUP_H_
#define AGBNP_SETUP_H_

#include <cstdint>
#include <string>
#include <vector>

namespace agbnp_b200_impl {

// float-literal constants of the reference, promoted to double exactly as its translation units see them
// (gaussvol/gaussvol.h:46-63, openmmapi/include/AGBNPForce.h:25-33, openmmapi/include/AGBNPUtils.h:124-126)
struct Constants {
    double kfc;          // (double) 2.2269859253f
    double volmina;      // (double)(0.01f*0.001f)
    double volminb;      // (double)(0.1f*0.001f)
    double min_gvol;     // (double) FLT_MIN
    double roffset;      // (double)(0.5f*0.1f)      AGBNP_RADIUS_INCREMENT
    double hb_radius;    // 1.4*(double)0.1f          AGBNP_HB_RADIUS
    double i4_maxa;      // 2.0
    int i4_nodes;        // 16
    int max_order;       // 8
    double dielectric_factor;  // 4.184*332/10*(-0.5)*(1 - 1/80)
    static Constants make();
};

// natural cubic spline tables of the switched pair descreening integral Q4(r; R_i, R_j), one per
// (screened radius type, screener radius type); node spacing h = maxa/(nodes-1)
struct I4Tables {
    int ntypes_screened = 0, ntypes_screener = 0, nodes = 0;
    double h = 0;
    std::vector<double> y, y2;              // [ti*ntypes_screener+tj][nodes]
    std::vector<int> type_screened;         // per atom
    std::vector<int> type_screener;         // per atom, -1 for hydrogens
    // per-interval packed form used on the device: (y_k, y_{k+1}, y2_k h^2/6, y2_{k+1} h^2/6)
    std::vector<float> packed;              // [table][nodes-1][4]
    void build(const std::vector<double>& radii, const std::vector<int>& ishydrogen, const Constants& c);
    double eval(double d, int ti, int tj) const;
    double evalderiv(double d, int ti, int tj) const;
};

// per-atom parameters as the kernel sees them + derived per-radius-type Gaussian constants
struct SystemParams {
    int n = 0;
    int version = 1;
    std::vector<double> radius, gamma, alpha, charge;   // gamma already zeroed for hydrogens
    std::vector<int> ishydrogen;
    double common_gamma = -1;
    // per-atom Gaussian constants (double, same expressions as gaussvol.cpp:131-132 / ReferenceAGBNPKernels.cpp:183,228):
    // a = KFC/r^2, v = 4 pi r^3/3 (0 for hydrogens), for the enlarged (L) and van der Waals (S) radii
    std::vector<double> aL, vL, aS, vS;
    // conservative level-2 pair-list filter: radii binned upward on a RC_BIN_WIDTH grid; rc2[bi*nbins+bj] bounds the
    // squared distance beyond which the enlarged-radius overlap volume is certainly below VOLMINA
    static constexpr double RC_BIN_WIDTH = 0.005;   // nm
    int nbins = 0;
    std::vector<int> rc_bin;                             // per atom
    std::vector<float> rc2;                              // [nbins*nbins]
    std::vector<float> rc2max;                           // [nbins] max over partners
    I4Tables i4;
    // returns "" or the reference's error message
    std::string init(int version, int n, const double* radius, const double* gamma, const double* alpha,
                     const double* charge, const unsigned char* ishydrogen, const Constants& c);
    std::string update(int n, const double* radius, const double* gamma, const double* alpha, const double* charge,
                       const unsigned char* ishydrogen);
};

// spatial ordering: heavy atoms first then hydrogens, each group in Morton order of 3-D cells
void morton_order(const float* xyz /*stride 4 or 3*/, int stride, const std::vector<int>& ishydrogen,
                  std::vector<int>& heavy_sorted, std::vector<int>& hydrogen_sorted);

// host-buffer marshalling of agbnp_b200_execute_host (compiled by the host compiler alone so that it can carry an AVX2
// clone next to the baseline one; chosen at load time)
void pack_positions(const double* pos /*[3n]*/, float* posq /*[4n], w = 0*/, int n);
void add_forces(const float* src, double* dst /* += */, int n3);
void set_forces(const float* src, double* dst /* = */, int n3);

} // namespace agbnp_b200_impl
#endif
