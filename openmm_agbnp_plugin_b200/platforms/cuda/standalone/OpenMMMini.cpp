// see openmm/OpenMMMini.h
#include "openmm/OpenMMMini.h"

namespace OpenMM {

static std::map<std::string, Platform*>& registry() { static std::map<std::string, Platform*> r; return r; }

Platform& Platform::getPlatformByName(const std::string& name) {
    std::map<std::string, Platform*>::iterator it = registry().find(name);
    if (it == registry().end()) throw OpenMMException("There is no registered Platform called \"" + name + "\"");
    return *it->second;
}
void Platform::registerPlatform(Platform* p) { registry()[p->getName()] = p; }

System::~System() { for (size_t i = 0; i < forces.size(); i++) delete forces[i]; }

ForceImpl& Force::getImplInContext(Context& context) { return context.getImpl().getImpl(this); }
ContextImpl& Force::getContextImpl(Context& context) { return context.getImpl(); }

ContextImpl::ContextImpl(Context& owner, const System& system, Platform& platform, int device)
    : owner(&owner), system(&system), platform(&platform) {
    positions.resize(system.getNumParticles());
    forces.resize(system.getNumParticles());
    data.positions = &positions; data.forces = &forces; data.device = device;
    for (int i = 0; i < system.getNumForces(); i++) {
        impls.push_back(system.getForce(i).createImpl());
        impls.back()->initialize(*this);
    }
}
ContextImpl::~ContextImpl() { for (size_t i = 0; i < impls.size(); i++) delete impls[i]; }

ForceImpl& ContextImpl::getImpl(const Force* f) {
    for (size_t i = 0; i < impls.size(); i++) if (&impls[i]->getOwner() == f) return *impls[i];
    throw OpenMMException("getImplInContext: the Force is not part of this Context");
}

double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy) {
    for (size_t i = 0; i < forces.size(); i++) forces[i] = Vec3();
    double e = 0.0;
    for (size_t i = 0; i < impls.size(); i++) e += impls[i]->calcForcesAndEnergy(*this, includeForces, includeEnergy, -1);
    return e;
}

} // namespace OpenMM
