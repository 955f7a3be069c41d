#ifndef ORACLE_SHIM_REALVEC_H_
#define ORACLE_SHIM_REALVEC_H_
#include "openmm/Vec3.h"
#include "openmm/reference/SimTKOpenMMRealType.h"
namespace OpenMM { typedef Vec3 RealVec; }
#endif
