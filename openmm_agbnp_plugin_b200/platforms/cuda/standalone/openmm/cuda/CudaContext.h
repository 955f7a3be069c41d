// Stand-in for the part of OpenMM's CudaContext (platforms/cuda/include/CudaContext.h) that a force plugin touches: the
// posq / force / energy buffers in the platform's own atom order and precision, the atom index map with its reorder
// listeners, the padded atom count and the stream.  Same method names and meanings as OpenMM 7.x, so that
// src/CudaAGBNPKernels.cpp compiles unchanged against the real headers (-DAGBNP_B200_WITH_OPENMM) and against this file.
// Stand-in only: setPositions / getForces / reorderAtoms, which in OpenMM belong to the integrator side of the context.
#ifndef AGBNP_B200_MOCK_CUDA_CONTEXT_H_
#define AGBNP_B200_MOCK_CUDA_CONTEXT_H_

#include <vector>

#include "openmm/OpenMMMini.h"
#include "openmm/cuda/CudaArray.h"

struct CUstream_st;
typedef CUstream_st* CUstream;

namespace OpenMM {

class CudaContext {
public:
    class ReorderListener {
    public:
        virtual void execute() = 0;
        virtual ~ReorderListener() {}
    };
    // precision: "single", "mixed" or "double" (CudaPlatform's Precision property)
    CudaContext(int numAtoms, int deviceIndex, const std::string& precision);
    ~CudaContext();
    void setAsCurrent();
    int getDeviceIndex() const { return deviceIndex; }
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    bool getUseDoublePrecision() const { return useDouble; }
    bool getUseMixedPrecision() const { return useMixed; }
    CudaArray& getPosq() { return posq; }                   // float4 (double4 in double precision) [paddedNumAtoms]
    CudaArray& getForce() { return force; }                 // long long [3*paddedNumAtoms], component-major, 2^32 fixed point
    CudaArray& getEnergyBuffer() { return energyBuffer; }   // float (double in mixed / double precision), summed by the context
    const std::vector<int>& getAtomIndex() const { return atomIndex; }      // atomIndex[i] = original index of the atom at position i
    void addReorderListener(ReorderListener* listener) { listeners.push_back(listener); }   // owned by the context
    CUstream getCurrentStream() { return 0; }
    // ---- stand-in only ----
    void setPositions(const std::vector<Vec3>& positions);
    void clearBuffers();
    void getForces(std::vector<Vec3>& forces);              // in original atom order
    double reduceEnergy();
    void reorderAtoms();                                    // pick another atom order, as the real context does every few hundred steps
private:
    void uploadPositions();
    int numAtoms, paddedNumAtoms, deviceIndex;
    bool useDouble, useMixed;
    CudaArray posq, force, energyBuffer;
    std::vector<int> atomIndex;
    std::vector<ReorderListener*> listeners;
    std::vector<Vec3> positions;
    int reorders;
};

} // namespace OpenMM
#endif
