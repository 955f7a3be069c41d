// see openmm/OpenMMMini.h
#include "openmm/OpenMMMini.h"
#include "openmm/cuda/CudaPlatform.h"

#include <cuda_runtime_api.h>

#include <algorithm>

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

ContextImpl::ContextImpl(Context& owner, const System& system, Platform& platform, int device, const std::string& precision)
    : owner(&owner), system(&system), platform(&platform), data(0) {
    positions.resize(system.getNumParticles());
    forces.resize(system.getNumParticles());
    data = new CudaPlatform::PlatformData(system.getNumParticles(), device, precision);
    for (int i = 0; i < system.getNumForces(); i++) {
        impls.push_back(system.getForce(i).createImpl());
        impls.back()->initialize(*this);
    }
}
ContextImpl::~ContextImpl() {
    for (size_t i = 0; i < impls.size(); i++) delete impls[i];
    delete static_cast<CudaPlatform::PlatformData*>(data);
}
CudaContext& ContextImpl::getCudaContext() { return *static_cast<CudaPlatform::PlatformData*>(data)->contexts[0]; }
void Context::setPositions(const std::vector<Vec3>& p) { impl->positions = p; impl->getCudaContext().setPositions(p); }

ForceImpl& ContextImpl::getImpl(const Force* f) {
    for (size_t i = 0; i < impls.size(); i++) if (&impls[i]->getOwner() == f) return *impls[i];
    throw OpenMMException("getImplInContext: the Force is not part of this Context");
}

double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy) {
    // as OpenMM's CUDA platform: clear the device buffers, let every force add to them, reduce
    CudaContext& cu = getCudaContext();
    cu.clearBuffers();
    double e = 0.0;
    for (size_t i = 0; i < impls.size(); i++) e += impls[i]->calcForcesAndEnergy(*this, includeForces, includeEnergy, -1);
    if (includeForces) cu.getForces(forces);
    if (includeEnergy) e += cu.reduceEnergy();
    return e;
}

// ---- CudaArray / CudaContext stand-ins ----
static void ck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw OpenMMException(std::string(what) + ": " + cudaGetErrorString(e));
}

CudaArray::~CudaArray() { if (ptr) cudaFree((void*) ptr); }
void CudaArray::initialize(size_t elements, size_t elementSize, const std::string& nm) {
    count = elements; elemSize = elementSize; name = nm;
    void* p = 0;
    ck(cudaMalloc(&p, count*elemSize), ("CudaArray " + nm).c_str());
    ptr = (CUdeviceptr_t) p;
    clear();
}
void CudaArray::upload(const void* d) { ck(cudaMemcpy((void*) ptr, d, count*elemSize, cudaMemcpyHostToDevice), "CudaArray::upload"); }
void CudaArray::download(void* d) const { ck(cudaMemcpy(d, (const void*) ptr, count*elemSize, cudaMemcpyDeviceToHost), "CudaArray::download"); }
void CudaArray::clear() { ck(cudaMemset((void*) ptr, 0, count*elemSize), "CudaArray::clear"); }

CudaContext::CudaContext(int numAtoms, int deviceIndex, const std::string& precision)
    : numAtoms(numAtoms), paddedNumAtoms((numAtoms+31)/32*32), deviceIndex(deviceIndex),
      useDouble(precision == "double"), useMixed(precision == "mixed"), reorders(0) {
    if (precision != "single" && precision != "mixed" && precision != "double") throw OpenMMException("Illegal value for Precision: " + precision);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw OpenMMException("No compatible CUDA device is available");
    setAsCurrent();
    posq.initialize(paddedNumAtoms, useDouble ? 4*sizeof(double) : 4*sizeof(float), "posq");
    force.initialize(3*(size_t) paddedNumAtoms, sizeof(long long), "force");
    energyBuffer.initialize(1024, (useDouble || useMixed) ? sizeof(double) : sizeof(float), "energyBuffer");
    atomIndex.resize(numAtoms);
    for (int i = 0; i < numAtoms; i++) atomIndex[i] = i;
    positions.resize(numAtoms);
}
CudaContext::~CudaContext() { for (size_t i = 0; i < listeners.size(); i++) delete listeners[i]; }
void CudaContext::setAsCurrent() { ck(cudaSetDevice(deviceIndex), "cudaSetDevice"); }
void CudaContext::setPositions(const std::vector<Vec3>& p) { positions = p; uploadPositions(); }
void CudaContext::uploadPositions() {
    setAsCurrent();
    if (useDouble) {
        std::vector<double> h(4*(size_t) paddedNumAtoms, 0.0);
        for (int i = 0; i < numAtoms; i++) for (int c = 0; c < 3; c++) h[4*(size_t) i+c] = positions[atomIndex[i]][c];
        posq.upload(h.data());
    } else {
        std::vector<float> h(4*(size_t) paddedNumAtoms, 0.f);
        for (int i = 0; i < numAtoms; i++) for (int c = 0; c < 3; c++) h[4*(size_t) i+c] = (float) positions[atomIndex[i]][c];
        posq.upload(h.data());
    }
}
void CudaContext::clearBuffers() { setAsCurrent(); force.clear(); energyBuffer.clear(); }
void CudaContext::getForces(std::vector<Vec3>& f) {
    setAsCurrent();
    ck(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    std::vector<long long> h(3*(size_t) paddedNumAtoms);
    force.download(h.data());
    const double scale = 1.0/4294967296.0;
    f.resize(numAtoms);
    for (int i = 0; i < numAtoms; i++)
        f[atomIndex[i]] = Vec3(scale*h[i], scale*h[(size_t) paddedNumAtoms+i], scale*h[2*(size_t) paddedNumAtoms+i]);
}
double CudaContext::reduceEnergy() {
    setAsCurrent();
    ck(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
    double e = 0.0;
    if (useDouble || useMixed) { std::vector<double> h(energyBuffer.getSize()); energyBuffer.download(h.data()); for (size_t i = 0; i < h.size(); i++) e += h[i]; }
    else { std::vector<float> h(energyBuffer.getSize()); energyBuffer.download(h.data()); for (size_t i = 0; i < h.size(); i++) e += h[i]; }
    return e;
}
void CudaContext::reorderAtoms() {
    // a different deterministic permutation every time (the real context sorts atoms along a space-filling curve)
    reorders++;
    std::vector<int> next(numAtoms);
    for (int i = 0; i < numAtoms; i++) next[i] = atomIndex[(int) (((long long) i*7919 + 13*reorders) % numAtoms)];
    std::vector<int> check(next);
    std::sort(check.begin(), check.end());
    for (int i = 0; i < numAtoms; i++) if (check[i] != i) { std::reverse(atomIndex.begin(), atomIndex.end()); next = atomIndex; break; }
    atomIndex = next;
    uploadPositions();
    for (size_t i = 0; i < listeners.size(); i++) listeners[i]->execute();
}

} // namespace OpenMM
