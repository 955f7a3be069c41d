// Stand-in for OpenMM's CudaPlatform (platforms/cuda/include/CudaPlatform.h): the platform object and the per-context
// PlatformData whose contexts[0] a kernel factory hands to its kernels.
#ifndef AGBNP_B200_MOCK_CUDA_PLATFORM_H_
#define AGBNP_B200_MOCK_CUDA_PLATFORM_H_

#include <vector>

#include "openmm/OpenMMMini.h"
#include "openmm/cuda/CudaContext.h"

namespace OpenMM {

class CudaPlatform : public Platform {
public:
    CudaPlatform() : Platform("CUDA") {}
    class PlatformData {
    public:
        PlatformData(int numAtoms, int deviceIndex, const std::string& precision) { contexts.push_back(new CudaContext(numAtoms, deviceIndex, precision)); }
        ~PlatformData() { for (size_t i = 0; i < contexts.size(); i++) delete contexts[i]; }
        std::vector<CudaContext*> contexts;
    };
};

} // namespace OpenMM
#endif
