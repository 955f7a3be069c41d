// Oracle shim: only PlatformData{positions, forces} is touched by the reference path
// (ReferenceAGBNPKernels.cpp:27-35).
#ifndef ORACLE_SHIM_REFERENCEPLATFORM_H_
#define ORACLE_SHIM_REFERENCEPLATFORM_H_
#include "openmm/Platform.h"
namespace OpenMM {
class ReferencePlatform : public Platform {
public:
    class PlatformData {
    public:
        void* positions;
        void* forces;
    };
};
}
#endif
