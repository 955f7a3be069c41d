// CudaCalcAGBNPForceKernel + factory + plugin registration (see CudaAGBNPKernels.h).
// One code path for OpenMM's CUDA platform and for the stand-in of this repository (platforms/cuda/standalone): positions
// are read from the context's posq buffer, forces are added to its 64-bit fixed-point force buffer and the energy to its
// energy buffer, all on the device and in the context's own atom order -- what platforms/opencl does with cl.getPosq() /
// cl.getForceBuffers() / cl.getEnergyBuffer() in the reference (OpenCLAGBNPKernels.cpp:602,1448,555).
#include "CudaAGBNPKernels.h"

#include <string>
#include <vector>

#include "openmm/cuda/CudaPlatform.h"

using namespace AGBNPPlugin;
using namespace OpenMM;

namespace {

struct ParamArrays {
    std::vector<double> radius, gamma, alpha, charge;
    std::vector<unsigned char> ishydrogen;
    explicit ParamArrays(const AGBNPForce& force) {
        const int n = force.getNumParticles();
        radius.resize(n); gamma.resize(n); alpha.resize(n); charge.resize(n); ishydrogen.resize(n);
        for (int i = 0; i < n; i++) {
            bool h;
            force.getParticleParameters(i, radius[i], gamma[i], alpha[i], charge[i], h);
            ishydrogen[i] = h ? 1 : 0;
        }
    }
};

void check(int rc, agbnp_b200* h) {
    if (rc != AGBNP_B200_OK) throw OpenMMException(agbnp_b200_last_error(h));
}

} // namespace

// CudaContext reorders its atoms from time to time (and tells its listeners): the library keeps its parameters by particle
// and only needs the new particle -> buffer position map
class CudaCalcAGBNPForceKernel::ReorderListener : public CudaContext::ReorderListener {
public:
    explicit ReorderListener(CudaCalcAGBNPForceKernel& owner) : owner(owner) {}
    void execute() { owner.syncDeviceLayout(); }
private:
    CudaCalcAGBNPForceKernel& owner;
};

CudaCalcAGBNPForceKernel::~CudaCalcAGBNPForceKernel() { agbnp_b200_destroy(handle); }

void CudaCalcAGBNPForceKernel::initialize(const System& system, const AGBNPForce& force) {
    (void) system;
    numParticles = force.getNumParticles();
    if (numParticles != cu.getNumAtoms()) throw OpenMMException("AGBNPForce must have exactly as many particles as the System it belongs to.");
    agbnp_b200_config cfg;
    agbnp_b200_default_config(&cfg);
    cfg.version = force.getVersion();
    cfg.nonbonded_method = (int) force.getNonbondedMethod();
    cfg.cutoff = force.getCutoffDistance();
    cfg.device = cu.getDeviceIndex();
    const ParamArrays p(force);
    check(agbnp_b200_create(&cfg, numParticles, p.radius.data(), p.gamma.data(), p.alpha.data(), p.charge.data(),
                            p.ishydrogen.data(), &handle), 0);
    syncDeviceLayout();
    cu.addReorderListener(new ReorderListener(*this));      // owned (and deleted) by the context
}

void CudaCalcAGBNPForceKernel::syncDeviceLayout() {
    if (!handle) return;
    agbnp_b200_device_layout lay;
    lay.atom_index = cu.getAtomIndex().data();
    lay.posq_is_double = cu.getUseDoublePrecision() ? 1 : 0;                                    // posq: double4 | float4
    lay.energy_is_float = (cu.getUseDoublePrecision() || cu.getUseMixedPrecision()) ? 0 : 1;    // energy buffer: "mixed" type
    check(agbnp_b200_set_device_layout(handle, &lay), handle);
}

double CudaCalcAGBNPForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    (void) context;
    cu.setAsCurrent();
    // Asynchronous: everything is enqueued on the context's stream.  A capacity overflow of an evaluation is repaired inside
    // the library but reported a few calls later (include/agbnp_b200.h, "Asynchronous use"): that step ran without the
    // AGBNP forces, so the simulation must not silently continue -- the exception carries the library's message.
    check(agbnp_b200_execute_device(handle, (const void*) cu.getPosq().getDevicePointer(), (void*) cu.getCurrentStream(),
                                    includeForces ? (void*) cu.getForce().getDevicePointer() : 0, 1 /* fixed point */, cu.getPaddedNumAtoms(),
                                    includeEnergy ? (double*) cu.getEnergyBuffer().getDevicePointer() : 0, 0), handle);
    return 0.0;                         // like the reference's OpenCL platform: the energy is in the context's buffer
}

void CudaCalcAGBNPForceKernel::copyParametersToContext(ContextImpl& context, const AGBNPForce& force) {
    (void) context;
    const ParamArrays p(force);
    check(agbnp_b200_set_params(handle, force.getNumParticles(), p.radius.data(), p.gamma.data(), p.alpha.data(), p.charge.data(),
                                p.ishydrogen.data()), handle);
}

KernelImpl* CudaAGBNPKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name != CalcAGBNPForceKernel::Name())
        throw OpenMMException("Tried to create kernel with illegal kernel name '" + name + "'");
    // same shape as the reference's OpenCL factory (OpenCLAGBNPKernelFactory.cpp:40-45)
    CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData())->contexts[0];
    return new CudaCalcAGBNPForceKernel(name, platform, cu);
}

// ---- OpenMM plugin entry points (the names are OpenMM's plugin ABI) ----
extern "C" void registerPlatforms() {}

extern "C" void registerKernelFactories() {
    try {
        Platform& platform = Platform::getPlatformByName("CUDA");
        platform.registerKernelFactory(CalcAGBNPForceKernel::Name(), new CudaAGBNPKernelFactory());
    } catch (const std::exception&) {
        // no CUDA platform in this OpenMM: nothing to register
    }
}

extern "C" void registerAGBNPCudaKernelFactories() {
    try {
        Platform::getPlatformByName("CUDA");
    } catch (const std::exception&) {
        Platform::registerPlatform(new CudaPlatform());
    }
    registerKernelFactories();
}
