#ifndef ORACLE_SHIM_FORCEIMPL_H_
#define ORACLE_SHIM_FORCEIMPL_H_
#include <map>
#include <string>
#include <vector>
namespace OpenMM {
class ContextImpl;
class ForceImpl {
public:
    virtual ~ForceImpl() {}
};
}
#endif
