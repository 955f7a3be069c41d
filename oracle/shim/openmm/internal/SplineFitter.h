// Oracle shim: restatement of the natural-cubic-spline arithmetic of OpenMM's SplineFitter
// (third-party, not under /root/reference; OpenMM "7.2.2 or later", unpinned).  Called by the reference at
// openmmapi/include/AGBNPUtils.h:104 (createNaturalSpline), :112 (evaluateSpline), :115 (evaluateSplineDerivative).
// Standard algorithm: tridiagonal solve for second derivatives with zero end conditions, binary search for the
// interval, cubic evaluation.  The evaluation half is also written out in the reference itself
// (platforms/opencl/src/kernels/AGBNPBornRadii.cl:58-73).  Pinned by platforms/reference/tests/v1.reference.
#ifndef ORACLE_SHIM_SPLINEFITTER_H_
#define ORACLE_SHIM_SPLINEFITTER_H_
#include <vector>
#include "openmm/OpenMMException.h"
namespace OpenMM {
class SplineFitter {
public:
    static void createNaturalSpline(const std::vector<double>& x, const std::vector<double>& y, std::vector<double>& y2) {
        int n = (int) x.size();
        if ((int) y.size() != n) throw OpenMMException("createNaturalSpline: x and y vectors must have same length");
        if (n < 2) throw OpenMMException("createNaturalSpline: the length of the input array must be at least 2");
        y2.assign(n, 0.0);
        if (n == 2) return;
        std::vector<double> lo(n, 0.0), di(n, 1.0), up(n, 0.0), rhs(n, 0.0), g(n, 0.0);
        for (int i = 1; i < n-1; i++) {
            lo[i] = x[i]-x[i-1];
            di[i] = 2.0*(x[i+1]-x[i-1]);
            up[i] = x[i+1]-x[i];
            rhs[i] = 6.0*((y[i+1]-y[i])/(x[i+1]-x[i]) - (y[i]-y[i-1])/(x[i]-x[i-1]));
        }
        // Thomas algorithm
        y2[0] = rhs[0]/di[0];
        double beta = di[0];
        for (int i = 1; i < n; i++) {
            g[i] = up[i-1]/beta;
            beta = di[i]-lo[i]*g[i];
            y2[i] = (rhs[i]-lo[i]*y2[i-1])/beta;
        }
        for (int i = n-2; i >= 0; i--) y2[i] -= g[i+1]*y2[i+1];
    }
    static double evaluateSpline(const std::vector<double>& x, const std::vector<double>& y, const std::vector<double>& y2, double t) {
        int lower, upper; locate(x, t, lower, upper);
        double dx = x[upper]-x[lower];
        double a = (x[upper]-t)/dx;
        double b = 1.0-a;
        return a*y[lower]+b*y[upper]+((a*a*a-a)*y2[lower]+(b*b*b-b)*y2[upper])*dx*dx/6.0;
    }
    static double evaluateSplineDerivative(const std::vector<double>& x, const std::vector<double>& y, const std::vector<double>& y2, double t) {
        int lower, upper; locate(x, t, lower, upper);
        double dx = x[upper]-x[lower];
        double a = (x[upper]-t)/dx;
        double b = 1.0-a;
        double dadx = -1.0/dx;
        return dadx*y[lower]-dadx*y[upper]+((1.0-3.0*a*a)*y2[lower]+(3.0*b*b-1.0)*y2[upper])*dx/6.0;
    }
private:
    static void locate(const std::vector<double>& x, double t, int& lower, int& upper) {
        int n = (int) x.size();
        if (t < x[0] || t > x[n-1]) throw OpenMMException("evaluateSpline: specified point is outside the range defined by the spline");
        lower = 0; upper = n-1;
        while (upper-lower > 1) {
            int middle = (upper+lower)/2;
            if (x[middle] > t) upper = middle; else lower = middle;
        }
    }
};
}
#endif
