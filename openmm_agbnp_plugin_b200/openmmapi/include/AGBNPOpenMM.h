// Selects the OpenMM headers: the real ones (-DAGBNP_B200_WITH_OPENMM, needs an OpenMM >= 7.2 install) or the minimal
// stand-in of platforms/cuda/standalone (this image has no OpenMM).
#ifndef AGBNP_OPENMM_SELECT_H_
#define AGBNP_OPENMM_SELECT_H_
#ifdef AGBNP_B200_WITH_OPENMM
#include "openmm/Context.h"
#include "openmm/Force.h"
#include "openmm/KernelImpl.h"
#include "openmm/OpenMMException.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/internal/ForceImpl.h"
#include "openmm/Kernel.h"
#include "openmm/KernelFactory.h"
#else
#include "openmm/OpenMMMini.h"
#endif
#endif
