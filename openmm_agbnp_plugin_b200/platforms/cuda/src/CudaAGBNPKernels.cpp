// CudaCalcAGBNPForceKernel + factory + plugin registration (see CudaAGBNPKernels.h).
#include "CudaAGBNPKernels.h"

#include <string>
#include <vector>

#ifdef AGBNP_B200_WITH_OPENMM
#include "openmm/cuda/CudaContext.h"
#include "openmm/cuda/CudaPlatform.h"
#endif

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

CudaCalcAGBNPForceKernel::~CudaCalcAGBNPForceKernel() { agbnp_b200_destroy(handle); }

void CudaCalcAGBNPForceKernel::initialize(const System& system, const AGBNPForce& force) {
    (void) system;
    numParticles = force.getNumParticles();
    agbnp_b200_config cfg;
    agbnp_b200_default_config(&cfg);
    cfg.version = force.getVersion();
    cfg.nonbonded_method = (int) force.getNonbondedMethod();
    cfg.cutoff = force.getCutoffDistance();
    cfg.device = device;
    const ParamArrays p(force);
    check(agbnp_b200_create(&cfg, numParticles, p.radius.data(), p.gamma.data(), p.alpha.data(), p.charge.data(),
                            p.ishydrogen.data(), &handle), 0);
}

double CudaCalcAGBNPForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
#ifdef AGBNP_B200_WITH_OPENMM
    // positions and forces stay on the GPU: posq float4 in, fixed-point force buffer and energy buffer out
    CudaContext& cu = *static_cast<CudaContext*>(platformContext);
    cu.setAsCurrent();
    double* d_energy = includeEnergy ? (double*) cu.getEnergyBuffer().getDevicePointer() : 0;   // mixed/double precision energy buffer
    check(agbnp_b200_execute_device(handle, (const void*) cu.getPosq().getDevicePointer(), (void*) cu.getCurrentStream(),
                                    includeForces ? (void*) cu.getForce().getDevicePointer() : 0, 1, cu.getPaddedNumAtoms(),
                                    d_energy, 0), handle);
    return 0.0;                         // like the OpenCL platform: the energy is in the buffer
#else
    // host arrays, the Reference platform's convention (ReferenceAGBNPKernels.cpp:27-35): energy returned, forces added
    HostPlatformData* data = static_cast<HostPlatformData*>(context.getPlatformData());
    std::vector<Vec3>& pos = *data->positions;
    std::vector<Vec3>& frc = *data->forces;
    posBuffer.resize(3*(size_t) numParticles);
    forceBuffer.assign(3*(size_t) numParticles, 0.0);
    for (int i = 0; i < numParticles; i++) for (int c = 0; c < 3; c++) posBuffer[3*(size_t) i+c] = pos[i][c];
    double energy = 0.0;
    check(agbnp_b200_execute_host(handle, posBuffer.data(), includeForces, includeEnergy, &energy, forceBuffer.data()), handle);
    if (includeForces)
        for (int i = 0; i < numParticles; i++) frc[i] += Vec3(forceBuffer[3*(size_t) i], forceBuffer[3*(size_t) i+1], forceBuffer[3*(size_t) i+2]);
    return energy;
#endif
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
#ifdef AGBNP_B200_WITH_OPENMM
    CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData())->contexts[0];
    return new CudaCalcAGBNPForceKernel(name, platform, &cu, cu.getDeviceIndex());
#else
    HostPlatformData* data = static_cast<HostPlatformData*>(context.getPlatformData());
    return new CudaCalcAGBNPForceKernel(name, platform, 0, data->device);
#endif
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
#ifdef AGBNP_B200_WITH_OPENMM
        Platform::registerPlatform(new CudaPlatform());
#else
        Platform::registerPlatform(new Platform("CUDA"));
#endif
    }
    registerKernelFactories();
}
