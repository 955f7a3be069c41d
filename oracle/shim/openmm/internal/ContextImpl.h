#ifndef ORACLE_SHIM_CONTEXTIMPL_H_
#define ORACLE_SHIM_CONTEXTIMPL_H_
#include "openmm/System.h"
#include "openmm/Platform.h"
namespace OpenMM {
class ContextImpl {
public:
    ContextImpl() : platformData(0) {}
    void* getPlatformData() { return platformData; }
    void setPlatformData(void* d) { platformData = d; }
    const System& getSystem() const { return system; }
    Platform& getPlatform() { return platform; }
private:
    void* platformData;
    System system;
    Platform platform;
};
}
#endif
