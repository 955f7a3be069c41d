// AGBNPForce / AGBNPForceImpl (see the headers for the reference lines each mirrors).
#include "AGBNPForce.h"
#include "AGBNPKernels.h"
#include "internal/AGBNPForceImpl.h"

using namespace AGBNPPlugin;
using OpenMM::OpenMMException;

// defaults of the reference constructor (AGBNPForce.cpp:15): NoCutoff, 1.0 nm, version 1; SOLVENT_RADIUS = 1.0*0.1f
AGBNPForce::AGBNPForce() : method(NoCutoff), cutoff(1.0), solvent_radius(1.0*0.1f), version(1) {}

int AGBNPForce::addParticle(double radius, double gamma, double vdw_alpha, double charge, bool ishydrogen) {
    Particle p = {radius, gamma, vdw_alpha, charge, ishydrogen};
    particles.push_back(p);
    return (int) particles.size()-1;
}

static void check_index(int index, size_t n) {
    if (index < 0 || index >= (int) n) throw OpenMMException("Index out of range");     // ASSERT_VALID_INDEX (AGBNPForce.cpp:25,64)
}

void AGBNPForce::getParticleParameters(int index, double& radius, double& gamma, double& vdw_alpha, double& charge, bool& ishydrogen) const {
    check_index(index, particles.size());
    const Particle& p = particles[index];
    radius = p.radius; gamma = p.gamma; vdw_alpha = p.alpha; charge = p.charge; ishydrogen = p.ishydrogen;
}

void AGBNPForce::setParticleParameters(int index, double radius, double gamma, double vdw_alpha, double charge, bool ishydrogen) {
    check_index(index, particles.size());
    Particle p = {radius, gamma, vdw_alpha, charge, ishydrogen};
    particles[index] = p;
}

void AGBNPForce::setVersion(int agbnp_version) {
    if (agbnp_version < 0 || agbnp_version > 2) throw OpenMMException("AGBNPForce::setVersion(): illegal version number");
    version = agbnp_version;
}

OpenMM::ForceImpl* AGBNPForce::createImpl() const { return new AGBNPForceImpl(*this); }

void AGBNPForce::updateParametersInContext(OpenMM::Context& context) {
    dynamic_cast<AGBNPForceImpl&>(getImplInContext(context)).updateParametersInContext(getContextImpl(context));
}

AGBNPForceImpl::~AGBNPForceImpl() {
#ifndef AGBNP_B200_WITH_OPENMM
    delete kernel.release();
#endif
}

void AGBNPForceImpl::initialize(OpenMM::ContextImpl& context) {
    kernel = context.getPlatform().createKernel(CalcAGBNPForceKernel::Name(), context);
    kernel.getAs<CalcAGBNPForceKernel>().initialize(context.getSystem(), owner);
}

double AGBNPForceImpl::calcForcesAndEnergy(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy, int groups) {
    if ((groups & (1 << owner.getForceGroup())) != 0) return kernel.getAs<CalcAGBNPForceKernel>().execute(context, includeForces, includeEnergy);
    return 0.0;
}

std::vector<std::string> AGBNPForceImpl::getKernelNames() { return std::vector<std::string>(1, CalcAGBNPForceKernel::Name()); }

void AGBNPForceImpl::updateParametersInContext(OpenMM::ContextImpl& context) {
    kernel.getAs<CalcAGBNPForceKernel>().copyParametersToContext(context, owner);
}
