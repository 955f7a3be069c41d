"""Python mirror of the reference's user-facing surface for the AGBNP1/GaussVol path.

Same names, argument meaning and error behaviour as the SWIG module `AGBNPplugin` (reference python/AGBNPPlugin.i:47-85,
openmmapi/include/AGBNPForce.h:39-155, openmmapi/src/AGBNPForce.cpp:15-78) and as the kernel interface
`CalcAGBNPForceKernel` (openmmapi/include/AGBNPKernels.h:19-47).  OpenMM itself is not available in this image, so a
minimal `Context` (positions + force accumulator) stands in for OpenMM's; all arithmetic happens in libagbnp_b200.so
through the C-ABI -- nothing here computes energies or forces.
"""
import ctypes as C

import numpy as np

from . import _lib


class OpenMMException(Exception):
    """Stands in for OpenMM::OpenMMException (same messages as the reference)."""


class AGBNPForce:
    # AGBNPForce::NonbondedMethod (AGBNPForce.h:44-59)
    NoCutoff = 0
    CutoffNonPeriodic = 1
    CutoffPeriodic = 2

    def __init__(self):
        # AGBNPForce.cpp:15: NoCutoff, cutoff 1.0 nm, version 1, solvent radius SOLVENT_RADIUS = 1.0*0.1f
        self._particles = []
        self._method = AGBNPForce.NoCutoff
        self._cutoff = 1.0
        self._version = 1
        self._solvent_radius = 1.0 * float(np.float32(0.1))
        self._impl = None

    def getNumParticles(self):
        return len(self._particles)

    def addParticle(self, radius, gamma, vdw_alpha, charge, ishydrogen):
        self._particles.append([float(radius), float(gamma), float(vdw_alpha), float(charge), bool(ishydrogen)])
        return len(self._particles) - 1

    def _check_index(self, index):
        if index < 0 or index >= len(self._particles):
            raise OpenMMException("Index out of range")   # ASSERT_VALID_INDEX (AGBNPForce.cpp:25,64)

    def setParticleParameters(self, index, radius, gamma, vdw_alpha, charge, ishydrogen):
        self._check_index(index)
        self._particles[index] = [float(radius), float(gamma), float(vdw_alpha), float(charge), bool(ishydrogen)]

    def getParticleParameters(self, index):
        self._check_index(index)
        r, g, a, q, h = self._particles[index]
        return r, g, a, q, h

    def getNonbondedMethod(self):
        return self._method

    def setNonbondedMethod(self, method):
        self._method = int(method)

    def getCutoffDistance(self):
        return self._cutoff

    def setCutoffDistance(self, distance):
        self._cutoff = float(distance)

    def getSolventRadius(self):
        return self._solvent_radius

    def setVersion(self, agbnp_version):
        # AGBNPForce.cpp:52-59
        if 0 <= int(agbnp_version) <= 2:
            self._version = int(agbnp_version)
        else:
            raise OpenMMException("AGBNPForce::setVersion(): illegal version number")

    def getVersion(self):
        return self._version

    def updateParametersInContext(self, context):
        # AGBNPForce.cpp:76-78 -> AGBNPForceImpl::updateParametersInContext -> kernel.copyParametersToContext
        kernel = context._kernel_for(self)
        kernel.copyParametersToContext(context, self)

    def _arrays(self):
        p = self._particles
        n = len(p)
        radius = np.array([x[0] for x in p], dtype=np.float64).reshape(n)
        gamma = np.array([x[1] for x in p], dtype=np.float64).reshape(n)
        alpha = np.array([x[2] for x in p], dtype=np.float64).reshape(n)
        charge = np.array([x[3] for x in p], dtype=np.float64).reshape(n)
        ish = np.array([1 if x[4] else 0 for x in p], dtype=np.uint8).reshape(n)
        return radius, gamma, alpha, charge, ish


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class CalcAGBNPForceKernel:
    """The CUDA-platform kernel object: forwards initialize/execute/copyParametersToContext to the C-ABI.

    Mirrors what platforms/cuda/CudaCalcAGBNPForceKernel (C++, see openmm_agbnp_plugin_b200/platforms/cuda) does when
    OpenMM is present."""

    @staticmethod
    def Name():
        return "CalcAGBNPForce"

    def __init__(self, name="CalcAGBNPForce", platform=None, device=0, shard_rank=0, shard_count=1, tree_reuse_interval=0):
        self.name = name
        self.platform = platform
        self.device = device
        self.shard_rank = shard_rank
        self.shard_count = shard_count
        # opt-in, NOT the reference's semantics: keep the overlap tree's topology for this many evaluations (MD)
        self.tree_reuse_interval = tree_reuse_interval
        self.handle = None

    def _err(self):
        L = _lib.lib()
        return (L.agbnp_b200_last_error(self.handle) or b"").decode()

    def initialize(self, system, force):
        L = _lib.lib()
        cfg = _lib.Config()
        L.agbnp_b200_default_config(C.byref(cfg))
        cfg.version = force.getVersion()
        cfg.nonbonded_method = force.getNonbondedMethod()
        cfg.cutoff = force.getCutoffDistance()
        cfg.device = self.device
        cfg.shard_rank = self.shard_rank
        cfg.shard_count = self.shard_count
        cfg.tree_reuse_interval = self.tree_reuse_interval
        radius, gamma, alpha, charge, ish = force._arrays()
        h = C.c_void_p()
        rc = L.agbnp_b200_create(C.byref(cfg), len(radius), _dp(radius), _dp(gamma), _dp(alpha), _dp(charge),
                                 ish.ctypes.data_as(C.POINTER(C.c_ubyte)), C.byref(h))
        if rc != _lib.OK:
            raise OpenMMException((L.agbnp_b200_last_error(None) or b"").decode())
        self.handle = h
        self.n = len(radius)

    def execute(self, context, includeForces=True, includeEnergy=True, assign=False):
        """Returns the potential energy (kJ/mol); forces are ADDED to context.forces (Reference-platform convention), or
        assigned if the caller says it would zero them first (AGBNP_B200_FORCES_ASSIGN)."""
        L = _lib.lib()
        pos = np.ascontiguousarray(context.positions, dtype=np.float64).reshape(-1)
        e = C.c_double(0.0)
        f = context.forces.reshape(-1)
        rc = L.agbnp_b200_execute_host(self.handle, _dp(pos), (2 if assign else 1) if includeForces else 0, int(includeEnergy), C.byref(e), _dp(f))
        if rc != _lib.OK:
            raise OpenMMException(self._err())
        return e.value

    def copyParametersToContext(self, context, force):
        L = _lib.lib()
        radius, gamma, alpha, charge, ish = force._arrays()
        rc = L.agbnp_b200_set_params(self.handle, len(radius), _dp(radius), _dp(gamma), _dp(alpha), _dp(charge),
                                     ish.ctypes.data_as(C.POINTER(C.c_ubyte)))
        if rc != _lib.OK:
            raise OpenMMException(self._err())

    # ---- diagnostics (agbnp_b200_get) ----
    def get(self, what):
        L = _lib.lib()
        sel = _lib.GET[what]
        if what in ("SCALARS", "WORK_COUNTERS", "STATS"):
            out = np.zeros(8)
        elif what in ("TREE_SIZE", "NEIGHBOR_COUNT"):
            out = np.zeros(1, dtype=np.int64)
        elif what == "TREE_TOPOLOGY":
            m = int(self.get("TREE_SIZE")[0])
            out = np.zeros((max(m, 1), 4), dtype=np.int32)
        elif what == "NEIGHBOR_PAIRS":
            m = int(self.get("NEIGHBOR_COUNT")[0])
            out = np.zeros((max(m, 1), 2), dtype=np.int32)
        else:
            out = np.zeros(self.n)
        rc = L.agbnp_b200_get(self.handle, sel, out.ctypes.data_as(C.c_void_p), out.nbytes)
        if rc != _lib.OK:
            raise OpenMMException(self._err())
        if what == "TREE_TOPOLOGY":
            return out[:int(self.get("TREE_SIZE")[0])]
        if what == "NEIGHBOR_PAIRS":
            return out[:int(self.get("NEIGHBOR_COUNT")[0])]
        return out

    def close(self):
        if self.handle is not None:
            _lib.lib().agbnp_b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """Minimal stand-in for an OpenMM Context holding one AGBNPForce: positions in, energy + forces out.

    Context(force, device=0) plays the role of AGBNPForce::createImpl + AGBNPForceImpl::initialize
    (openmmapi/src/AGBNPForceImpl.cpp:27-30): it creates the platform kernel by name and initializes it."""

    def __init__(self, force, device=0, shard_rank=0, shard_count=1, tree_reuse_interval=0):
        self.force = force
        self.kernel = CalcAGBNPForceKernel(CalcAGBNPForceKernel.Name(), None, device, shard_rank, shard_count, tree_reuse_interval)
        self.kernel.initialize(None, force)
        n = force.getNumParticles()
        self.positions = np.zeros((n, 3))
        self.forces = np.zeros((n, 3))
        self.energy = 0.0

    def _kernel_for(self, force):
        return self.kernel

    def setPositions(self, positions):
        self.positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)

    def calcForcesAndEnergy(self, includeForces=True, includeEnergy=True):
        # AGBNPForceImpl::calcForcesAndEnergy (AGBNPForceImpl.cpp:32-36) behind ContextImpl's "zero the forces, then let every
        # force add its own": with one force in the context that is an assignment
        if not includeForces:
            self.forces[:] = 0.0
        self.energy = self.kernel.execute(self, includeForces, includeEnergy, assign=True)
        return self.energy

    def getPotentialEnergy(self):
        return self.energy

    def getForces(self):
        return self.forces
