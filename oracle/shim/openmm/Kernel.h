#ifndef ORACLE_SHIM_KERNEL_H_
#define ORACLE_SHIM_KERNEL_H_
#include "openmm/KernelImpl.h"
namespace OpenMM {
class Kernel {
public:
    Kernel() : impl(0) {}
    Kernel(KernelImpl* impl) : impl(impl) {}
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
private:
    KernelImpl* impl;
};
}
#endif
