#ifndef ORACLE_SHIM_FORCE_H_
#define ORACLE_SHIM_FORCE_H_
#include "openmm/Context.h"
namespace OpenMM {
class ForceImpl;
class ContextImpl;
class Force {
public:
    Force() : forceGroup(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return forceGroup; }
protected:
    virtual ForceImpl* createImpl() const = 0;
    ForceImpl& getImplInContext(Context& context);
    ContextImpl& getContextImpl(Context& context);
private:
    int forceGroup;
};
}
#endif
