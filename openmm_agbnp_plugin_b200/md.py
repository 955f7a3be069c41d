"""Minimal velocity-Verlet driver with the AGBNP force as the ONLY force, for the reference's end-to-end checks where
OpenMM is absent: `example/test_agbnp.py:55-64` integrates with Verlet and prints the total energy every 10
steps to check ENERGY CONSERVATION, `example/*_benchmark.py` time MD steps.  Positions, velocities and forces stay on the
GPU; each step is one asynchronous agbnp_b200_execute_device call (the CUDA-platform calling convention) plus three
element-wise torch updates.  Nothing here computes AGBNP energies or forces.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .AGBNPplugin import CalcAGBNPForceKernel, OpenMMException


class VerletNVE:
    def __init__(self, force, positions_nm, masses_amu, dt_ps=0.001, device=0, restraint_k=0.0, tree_reuse_interval=0):
        """restraint_k (kJ/mol/nm^2): optional harmonic tether of every atom to its initial position.  AGBNP alone has no
        bonded or repulsive terms, so an untethered solute collapses; the tether stands in for the rest of the force field
        (a torch expression, not part of the library) when the energy-conservation check needs a stable trajectory."""
        self.kernel = CalcAGBNPForceKernel(CalcAGBNPForceKernel.Name(), None, device, tree_reuse_interval=tree_reuse_interval)
        self.kernel.initialize(None, force)
        self.dev = torch.device("cuda", device)
        n = len(positions_nm)
        self.n = n
        self.posq = torch.zeros((n, 4), dtype=torch.float32, device=self.dev)
        self.posq[:, :3] = torch.as_tensor(np.asarray(positions_nm, dtype=np.float32), device=self.dev)
        self.vel = torch.zeros((n, 3), dtype=torch.float32, device=self.dev)
        self.frc = torch.zeros((n, 3), dtype=torch.float32, device=self.dev)
        self.inv_m = (1.0/torch.as_tensor(np.asarray(masses_amu, dtype=np.float32), device=self.dev)).unsqueeze(1)
        self.mass = 1.0/self.inv_m
        self.dt = float(dt_ps)
        self.k_res = float(restraint_k)
        self.x0 = self.posq[:, :3].clone()
        self.e_dev = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.steps = 0
        self._force(sync=True)

    def _force(self, sync=False):
        L = _lib.lib()
        self.frc.zero_()
        self.e_dev.zero_()
        st = torch.cuda.current_stream(self.dev).cuda_stream
        he = C.c_double(0.0)
        rc = L.agbnp_b200_execute_device(self.kernel.handle, C.c_void_p(self.posq.data_ptr()), C.c_void_p(st),
                                         C.c_void_p(self.frc.data_ptr()), 0, self.n, C.c_void_p(self.e_dev.data_ptr()),
                                         C.byref(he) if sync else None)
        if rc != _lib.OK:
            raise OpenMMException(self.kernel._err())
        if self.k_res:
            self.frc -= self.k_res*(self.posq[:, :3]-self.x0)

    def step(self, nsteps=1):
        """velocity Verlet; units nm, ps, amu, kJ/mol (force kJ/mol/nm -> acceleration nm/ps^2 = F/m)."""
        dt = self.dt
        for _ in range(nsteps):
            self.vel += (0.5*dt)*self.frc*self.inv_m
            self.posq[:, :3] += dt*self.vel
            self._force()
            self.vel += (0.5*dt)*self.frc*self.inv_m
            self.steps += 1

    def energies(self):
        """(potential, kinetic) in kJ/mol of the current state; synchronises."""
        L = _lib.lib()
        st = torch.cuda.current_stream(self.dev).cuda_stream
        if L.agbnp_b200_synchronize(self.kernel.handle, C.c_void_p(st)) != _lib.OK:
            raise OpenMMException(self.kernel._err())
        pe = float(self.e_dev.item())
        if self.k_res:
            pe += float(0.5*self.k_res*((self.posq[:, :3]-self.x0).double()**2).sum().item())
        ke = float(0.5*(self.mass*self.vel.double()**2).sum().item())
        return pe, ke

    def list_stats(self):
        out = np.zeros(4)
        _lib.lib().agbnp_b200_get(self.kernel.handle, _lib.GET["LIST_STATS"], out.ctypes.data_as(C.c_void_p), out.nbytes)
        return out

    def close(self):
        self.kernel.close()


_md_lib = None


def _md():
    """lib/libagbnp_md.so: the fused integrator kernel (csrc/agbnp_md.cu).  Fails loudly if it has not been built."""
    global _md_lib
    if _md_lib is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libagbnp_md.so")
        if not os.path.exists(path):
            raise RuntimeError("libagbnp_md.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.agbnp_md_langevin_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_float,
                                             C.c_float, C.c_float, C.c_ulonglong, C.c_ulonglong, C.c_void_p, C.c_void_p]
        _md_lib = L
    return _md_lib


KB = 0.0083144626      # kJ/mol/K


class LangevinMD:
    """The protocol of the reference's MD benchmarks (example/hivrt_benchmark.py:17-33: LangevinIntegrator(300 K, 1/ps, 1 fs),
    simulation.step(10000)) with the AGBNP force as the only force: per step ONE asynchronous AGBNP evaluation
    (agbnp_b200_execute_device, the CUDA-platform calling convention) and ONE fused integrator kernel; nothing touches the
    host.  `restraint_k` tethers every atom to its start position -- AGBNP alone has no bonded or repulsive terms, the tether
    stands in for the rest of the force field so that the structure (hence the overlap tree and the neighbor lists) behaves
    as in a real simulation: 20 000 kJ/mol/nm^2 gives the 0.01 nm thermal amplitude of bonded atoms at 300 K (a soft tether
    lets the spheres interpenetrate -- AGBNP has no repulsion -- and the overlap trees grow far beyond a protein's).

    Asynchronous evaluations that overflowed a capacity are reported a few steps later by the library (it has grown the
    capacity by then): such a step ran without AGBNP forces.  The driver counts them in `dropped` and goes on -- the
    thermostat absorbs the glitch -- so a caller that needs a clean trajectory checks `dropped == 0` (bench.py does)."""

    def __init__(self, force, positions_nm, masses_amu, temperature=300.0, friction_per_ps=1.0, dt_ps=0.001, device=0,
                 restraint_k=20000.0, seed=2026, tree_reuse_interval=0):
        self.kernel = CalcAGBNPForceKernel(CalcAGBNPForceKernel.Name(), None, device, tree_reuse_interval=tree_reuse_interval)
        self.kernel.initialize(None, force)
        self.dev = torch.device("cuda", device)
        n = len(positions_nm)
        self.n = n
        self.posq = torch.zeros((n, 4), dtype=torch.float32, device=self.dev)
        self.posq[:, :3] = torch.as_tensor(np.asarray(positions_nm, dtype=np.float32), device=self.dev)
        self.x0 = self.posq.clone()
        self.vel = torch.zeros((n, 3), dtype=torch.float32, device=self.dev)
        self.frc = torch.zeros((n, 3), dtype=torch.float32, device=self.dev)
        m = np.asarray(masses_amu, dtype=np.float64)
        self.inv_m = torch.as_tensor((1.0/m).astype(np.float32), device=self.dev)
        self.ke2 = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.e_dev = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.kT = KB*float(temperature)
        self.friction, self.dt, self.k_res, self.seed = float(friction_per_ps), float(dt_ps), float(restraint_k), int(seed)
        self.steps = 0
        self.dropped = 0
        # Maxwell-Boltzmann start (the reference scripts read velocities from the .dms file)
        g = torch.Generator(device=self.dev)
        g.manual_seed(seed)
        self.vel.copy_(torch.randn((n, 3), generator=g, device=self.dev)*torch.sqrt(self.kT*self.inv_m).unsqueeze(1))
        self._force(sync=True)

    def _force(self, sync=False):
        L = _lib.lib()
        st = torch.cuda.current_stream(self.dev).cuda_stream
        he = C.c_double(0.0)
        rc = L.agbnp_b200_execute_device(self.kernel.handle, C.c_void_p(self.posq.data_ptr()), C.c_void_p(st),
                                         C.c_void_p(self.frc.data_ptr()), 0, self.n, C.c_void_p(self.e_dev.data_ptr()),
                                         C.byref(he) if sync else None)
        if rc == _lib.ERR_CAPACITY and not sync:
            self.dropped += 1               # an EARLIER evaluation was not delivered; this one was enqueued as usual
        elif rc != _lib.OK:
            raise OpenMMException(self.kernel._err())

    def step(self, nsteps=1, measure_ke=False):
        st = torch.cuda.current_stream(self.dev).cuda_stream
        M = _md()
        for _ in range(nsteps):
            rc = M.agbnp_md_langevin_step(C.c_void_p(self.posq.data_ptr()), C.c_void_p(self.vel.data_ptr()), C.c_void_p(self.frc.data_ptr()),
                                          C.c_void_p(self.inv_m.data_ptr()), C.c_void_p(self.x0.data_ptr()) if self.k_res else None,
                                          self.k_res, self.n, self.dt, self.friction, self.kT, self.seed, self.steps,
                                          C.c_void_p(self.ke2.data_ptr()) if measure_ke else None, C.c_void_p(st))
            if rc != 0:
                raise OpenMMException("agbnp_md_langevin_step failed")
            self._force()
            self.steps += 1

    def synchronize(self):
        st = torch.cuda.current_stream(self.dev).cuda_stream
        rc = _lib.lib().agbnp_b200_synchronize(self.kernel.handle, C.c_void_p(st))
        if rc == _lib.ERR_CAPACITY:
            self.dropped += 1
        elif rc != _lib.OK:
            raise OpenMMException(self.kernel._err())

    def temperature(self):
        """instantaneous kinetic temperature (K)"""
        self.synchronize()
        ke2 = float(((self.vel.double()**2).sum(dim=1)/self.inv_m.double()).sum().item())
        return ke2/(3.0*self.n*KB)

    def stats(self):
        out = np.zeros(8)
        _lib.lib().agbnp_b200_get(self.kernel.handle, _lib.GET["STATS"], out.ctypes.data_as(C.c_void_p), out.nbytes)
        return out

    def list_stats(self):
        out = np.zeros(4)
        _lib.lib().agbnp_b200_get(self.kernel.handle, _lib.GET["LIST_STATS"], out.ctypes.data_as(C.c_void_p), out.nbytes)
        return out

    def close(self):
        self.kernel.close()
