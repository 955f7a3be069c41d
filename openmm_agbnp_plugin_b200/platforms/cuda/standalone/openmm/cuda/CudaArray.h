// Stand-in for OpenMM's CudaArray (platforms/cuda/include/CudaArray.h): a typed device allocation whose
// getDevicePointer() is what plugin kernels pass to their launches.  Only what the AGBNP plugin uses.
#ifndef AGBNP_B200_MOCK_CUDA_ARRAY_H_
#define AGBNP_B200_MOCK_CUDA_ARRAY_H_

#include <cstddef>
#include <string>

namespace OpenMM {

typedef unsigned long long CUdeviceptr_t;      // CUdeviceptr of the driver API

class CudaArray {
public:
    CudaArray() : ptr(0), count(0), elemSize(0) {}
    ~CudaArray();
    void initialize(size_t elements, size_t elementSize, const std::string& name);
    CUdeviceptr_t& getDevicePointer() { return ptr; }
    size_t getSize() const { return count; }
    int getElementSize() const { return (int) elemSize; }
    void upload(const void* data);
    void download(void* data) const;
    void clear();
private:
    CUdeviceptr_t ptr;
    size_t count, elemSize;
    std::string name;
};

} // namespace OpenMM
#endif
