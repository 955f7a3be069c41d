// CudaCalcAGBNPForceKernel -- the B200 (sm_100a) implementation of CalcAGBNPForceKernel.  All arithmetic lives behind
// the C-ABI of libagbnp_b200.so (include/agbnp_b200.h); this class only moves parameters in and forwards execute.
// It takes the place of the reference's OpenCLCalcAGBNPForceKernel (platforms/opencl/src/OpenCLAGBNPKernels.h).
#ifndef AGBNP_B200_CUDA_KERNELS_H_
#define AGBNP_B200_CUDA_KERNELS_H_

#include "AGBNPKernels.h"
#include "agbnp_b200.h"

namespace AGBNPPlugin {

class CudaCalcAGBNPForceKernel : public CalcAGBNPForceKernel {
public:
    CudaCalcAGBNPForceKernel(std::string name, const OpenMM::Platform& platform, void* platformContext, int device)
        : CalcAGBNPForceKernel(name, platform), platformContext(platformContext), device(device), handle(0), numParticles(0) {}
    ~CudaCalcAGBNPForceKernel();
    void initialize(const OpenMM::System& system, const AGBNPForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
    void copyParametersToContext(OpenMM::ContextImpl& context, const AGBNPForce& force);
    agbnp_b200* getHandle() { return handle; }
private:
    void* platformContext;      // CudaContext* with OpenMM, unused in the standalone build
    int device;
    agbnp_b200* handle;
    int numParticles;
    std::vector<double> posBuffer, forceBuffer;
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
