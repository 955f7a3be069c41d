// The abstract kernel every platform implements for AGBNPForce (reference openmmapi/include/AGBNPKernels.h:19-47): the
// kernel name "CalcAGBNPForce" and the three virtuals are the contract between AGBNPForceImpl and a platform plugin.
#ifndef AGBNP_B200_AGBNPKERNELS_H_
#define AGBNP_B200_AGBNPKERNELS_H_

#include <string>

#include "AGBNPForce.h"

namespace AGBNPPlugin {

class CalcAGBNPForceKernel : public OpenMM::KernelImpl {
public:
    static std::string Name() { return "CalcAGBNPForce"; }
    CalcAGBNPForceKernel(std::string name, const OpenMM::Platform& platform) : OpenMM::KernelImpl(name, platform) {}
    virtual void initialize(const OpenMM::System& system, const AGBNPForce& force) = 0;
    // returns the potential energy; forces are accumulated into the platform's force buffer
    virtual double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy) = 0;
    virtual void copyParametersToContext(OpenMM::ContextImpl& context, const AGBNPForce& force) = 0;
};

} // namespace AGBNPPlugin
#endif
