// AGBNPForceImpl -- binds an AGBNPForce to the platform kernel inside a Context (reference
// openmmapi/include/internal/AGBNPForceImpl.h:17-45, openmmapi/src/AGBNPForceImpl.cpp:27-46).
#ifndef AGBNP_B200_AGBNPFORCEIMPL_H_
#define AGBNP_B200_AGBNPFORCEIMPL_H_

#include <map>
#include <string>
#include <vector>

#include "AGBNPForce.h"

namespace AGBNPPlugin {

class AGBNPForceImpl : public OpenMM::ForceImpl {
public:
    explicit AGBNPForceImpl(const AGBNPForce& owner) : owner(owner) {}
    ~AGBNPForceImpl();
    void initialize(OpenMM::ContextImpl& context);
    const AGBNPForce& getOwner() const { return owner; }
    void updateContextState(OpenMM::ContextImpl&) {}
    double calcForcesAndEnergy(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy, int groups);
    std::map<std::string, double> getDefaultParameters() { return std::map<std::string, double>(); }
    std::vector<std::string> getKernelNames();
    void updateParametersInContext(OpenMM::ContextImpl& context);
private:
    const AGBNPForce& owner;
    OpenMM::Kernel kernel;
};

} // namespace AGBNPPlugin
#endif
