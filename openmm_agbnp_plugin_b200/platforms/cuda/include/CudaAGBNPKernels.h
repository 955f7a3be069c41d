// CudaCalcAGBNPForceKernel -- the B200 (sm_100a) implementation of CalcAGBNPForceKernel.  All arithmetic lives behind
// the C-ABI of libagbnp_b200.so (include/agbnp_b200.h); this class only moves parameters in and forwards execute.
// It takes the place of the reference's OpenCLCalcAGBNPForceKernel (platforms/opencl/src/OpenCLAGBNPKernels.h).
#ifndef AGBNP_B200_CUDA_KERNELS_H_
#define AGBNP_B200_CUDA_KERNELS_H_

#include "AGBNPKernels.h"
#include "agbnp_b200.h"
#include "openmm/cuda/CudaContext.h"     // OpenMM's, or the stand-in under standalone/ (same interface)

namespace AGBNPPlugin {

class CudaCalcAGBNPForceKernel : public CalcAGBNPForceKernel {
public:
    CudaCalcAGBNPForceKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu)
        : CalcAGBNPForceKernel(name, platform), cu(cu), handle(0), numParticles(0) {}
    ~CudaCalcAGBNPForceKernel();
    void initialize(const OpenMM::System& system, const AGBNPForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
    void copyParametersToContext(OpenMM::ContextImpl& context, const AGBNPForce& force);
    agbnp_b200* getHandle() { return handle; }
    // tell the library where each particle lives in the context's buffers (CudaContext::getAtomIndex) and which precision
    // they have; called at initialize and by the reorder listener
    void syncDeviceLayout();
private:
    class ReorderListener;
    OpenMM::CudaContext& cu;
    agbnp_b200* handle;
    int numParticles;
};

class CudaAGBNPKernelFactory : public OpenMM::KernelFactory {
public:
    OpenMM::KernelImpl* createKernelImpl(std::string name, const OpenMM::Platform& platform, OpenMM::ContextImpl& context) const;
};

} // namespace AGBNPPlugin

extern "C" void registerPlatforms();
extern "C" void registerKernelFactories();
extern "C" void registerAGBNPCudaKernelFactories();
#endif
