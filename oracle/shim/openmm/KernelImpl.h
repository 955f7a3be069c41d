#ifndef ORACLE_SHIM_KERNELIMPL_H_
#define ORACLE_SHIM_KERNELIMPL_H_
#include <string>
#include "openmm/Platform.h"
namespace OpenMM {
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    std::string name;
    const Platform* platform;
};
}
#endif
