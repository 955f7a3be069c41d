#ifndef ORACLE_SHIM_CONTEXT_H_
#define ORACLE_SHIM_CONTEXT_H_
#include "openmm/System.h"
#include "openmm/Platform.h"
namespace OpenMM { class Context {}; }
#endif
