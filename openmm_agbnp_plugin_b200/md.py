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

    def close(self):
        self.kernel.close()
