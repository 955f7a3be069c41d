// TestCudaAGBNPForce [version] [precision: single|mixed|double] < system.dat
// The CUDA-platform twin of the reference's platforms/reference/tests/TestReferenceAGBNPForce.cpp: same stdin format
// (N, then per line: id x y z radius[A] charge gamma[kcal/mol/A^2] ishydrogen), same unit conversions and alpha rule
// (:47-70), same "Energy:" output, plus the finite-difference lines of v0.reference / v1.reference (atom 121, +2e-3 nm in y).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "AGBNPForce.h"
#include "CudaAGBNPKernels.h"

using namespace AGBNPPlugin;
using namespace OpenMM;

int main(int argc, char** argv) {
    try {
        registerAGBNPCudaKernelFactories();
        const int version = argc > 1 ? std::atoi(argv[1]) : 1;
        System system;
        AGBNPForce* force = new AGBNPForce();
        force->setVersion(version);
        system.addForce(force);
        int n = 0;
        std::cin >> n;
        const double ang2nm = 0.1, kcal2kj = 4.184;
        const double sigmaw = 3.15365*ang2nm, epsilonw = 0.155*kcal2kj, rho = 0.033428/std::pow(ang2nm, 3), epsilon_lj = 0.155*kcal2kj;
        std::vector<Vec3> positions;
        for (int i = 0; i < n; i++) {
            double id, x, y, z, radius, charge, gamma;
            int ih;
            std::cin >> id >> x >> y >> z >> radius >> charge >> gamma >> ih;
            system.addParticle(1.0);
            positions.push_back(Vec3(x, y, z)*ang2nm);
            radius *= ang2nm;
            gamma *= kcal2kj/(ang2nm*ang2nm);
            const double sij = std::sqrt(sigmaw*2.0*radius), eij = std::sqrt(epsilonw*epsilon_lj);
            const double alpha = -16.0*M_PI*rho*eij*std::pow(sij, 6)/3.0;
            force->addParticle(radius, gamma, alpha, charge, ih > 0);
        }
        const std::string precision = argc > 2 ? argv[2] : "single";
        Context context(system, Platform::getPlatformByName("CUDA"), 0, precision);
        context.setPositions(positions);
        const double e1 = context.getPotentialEnergy();
        const std::vector<Vec3> forces = context.getForces();
        std::cout << "Energy: " << e1 << std::endl;
#ifndef AGBNP_B200_WITH_OPENMM
        // The CUDA platform reorders its atoms between steps (CudaContext::reorderAtoms -> ReorderListener): the kernel must
        // follow the new buffer order, and no particle's force may change.  (Stand-in only: with OpenMM the context decides.)
        for (int round = 0; round < 2; round++) {
            context.getImpl().getCudaContext().reorderAtoms();
            const double er = context.getPotentialEnergy();
            const std::vector<Vec3>& fr = context.getForces();
            double dmax = 0.0, fmax = 0.0;
            for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) {
                dmax = std::max(dmax, std::fabs(fr[i][c]-forces[i][c])); fmax = std::max(fmax, std::fabs(forces[i][c]));
            }
            std::cout << "Reorder relative energy change: " << std::fabs(er-e1)/std::fabs(e1) << std::endl;
            std::cout << "Reorder relative force change: " << dmax/fmax << std::endl;
        }
#endif
        const int pmove = 121, direction = 1;
        if (n > pmove) {
            const double offset = 2.e-3;
            positions[pmove][direction] += offset;
            context.setPositions(positions);
            const double e2 = context.getPotentialEnergy();
            std::cout << "Energy: " << e2 << std::endl;
            std::cout << "Energy Change: " << e2-e1 << std::endl;
            std::cout << "Energy Change from Gradient: " << -forces[pmove][direction]*offset << std::endl;
        }
        // updateParametersInContext: doubling every charge must change the energy of AGBNP1 and is accepted;
        // changing a radius is rejected with the reference's message
        if (version == 1) {
            for (int i = 0; i < n; i++) {
                double r, g, a, q; bool h;
                force->getParticleParameters(i, r, g, a, q, h);
                force->setParticleParameters(i, r, g, a, 2.0*q, h);
            }
            force->updateParametersInContext(context);
            std::cout << "Energy after charge update: " << context.getPotentialEnergy() << std::endl;
        }
    } catch (const std::exception& e) {
        std::cout << "exception: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
