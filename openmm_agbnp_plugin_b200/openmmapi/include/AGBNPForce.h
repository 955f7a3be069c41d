// AGBNPForce -- the user-facing force class of the AGBNP plugin, same public surface as the reference's
// openmmapi/include/AGBNPForce.h:39-155 (method names, argument order and units are the plugin's public API and the SWIG
// module python/AGBNPPlugin.i:47-85 binds exactly these).  Units: nm, kJ/mol, e (reference README.md:97-103).
#ifndef AGBNP_B200_AGBNPFORCE_H_
#define AGBNP_B200_AGBNPFORCE_H_

#include <vector>

#include "AGBNPOpenMM.h"

namespace AGBNPPlugin {

class AGBNPForce : public OpenMM::Force {
public:
    // values as in the reference (AGBNPForce.h:44-59)
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2 };

    AGBNPForce();
    int getNumParticles() const { return (int) particles.size(); }
    // radius (nm), gamma (kJ/mol/nm^2), vdw_alpha (kJ nm^3/mol), charge (e); returns the particle index
    int addParticle(double radius, double gamma, double vdw_alpha, double charge, bool ishydrogen);
    void getParticleParameters(int index, double& radius, double& gamma, double& vdw_alpha, double& charge, bool& ishydrogen) const;
    void setParticleParameters(int index, double radius, double gamma, double vdw_alpha, double charge, bool ishydrogen);
    NonbondedMethod getNonbondedMethod() const { return method; }
    void setNonbondedMethod(NonbondedMethod m) { method = m; }
    double getCutoffDistance() const { return cutoff; }
    void setCutoffDistance(double distance) { cutoff = distance; }
    double getSolventRadius() const { return solvent_radius; }
    // only gamma, alpha and charge may have changed (ReferenceAGBNPKernels.cpp:1796-1815)
    void updateParametersInContext(OpenMM::Context& context);
    bool usesPeriodicBoundaryConditions() const { return false; }
    void setVersion(int agbnp_version);          // 0 GVolSA, 1 AGBNP1, 2 AGBNP2; anything else throws
    unsigned int getVersion() const { return (unsigned int) version; }
protected:
    OpenMM::ForceImpl* createImpl() const;
private:
    struct Particle { double radius, gamma, alpha, charge; bool ishydrogen; };
    std::vector<Particle> particles;
    NonbondedMethod method;
    double cutoff, solvent_radius;
    int version;
};

} // namespace AGBNPPlugin
#endif
