#ifndef ORACLE_SHIM_ASSERTIONUTILITIES_H_
#define ORACLE_SHIM_ASSERTIONUTILITIES_H_
#include "openmm/OpenMMException.h"
#define ASSERT_VALID_INDEX(index, vector) \
    { if ((index) < 0 || (index) >= (int) (vector).size()) throw OpenMM::OpenMMException("Index out of range"); }
#endif
