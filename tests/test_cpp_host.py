"""The C++ host layer above the C-ABI (openmm_agbnp_plugin_b200/openmmapi + platforms/cuda): AGBNPForce -> AGBNPForceImpl
-> CudaCalcAGBNPForceKernel, registered through the reference's plugin entry points.  The test program is the CUDA twin
of the reference's TestReferenceAGBNPForce (same stdin format, same output lines), so the GPU test reads like the
reference's own: feed gaussvol.dat, compare with v0.reference / v1.reference."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
PKG = os.path.join(ROOT, "openmm_agbnp_plugin_b200")
EXE = os.path.join(PKG, "platforms", "cuda", "tests", "TestCudaAGBNPForce")
PLUGIN = os.path.join(PKG, "lib", "libAGBNPPluginCUDA.so")


def _gaussvol_dat():
    """The reference's test input format (TestReferenceAGBNPForce.cpp:45-58) regenerated from the committed fixture:
    N, then id x y z radius[A] charge gamma[kcal/mol/A^2] ishydrogen per line."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "gaussvol.npz"))
    lines = ["%d" % len(d["radius"])]
    for i in range(len(d["radius"])):
        x, y, z = d["pos"][i] * 10.0
        lines.append("%d %.17g %.17g %.17g %.17g %.17g %.17g %d" % (i, x, y, z, d["radius"][i] * 10.0, d["charge"][i],
                                                                    d["gamma"][i] * 0.01 / 4.184, int(d["ishydrogen"][i])))
    return "\n".join(lines) + "\n"


def _sig(x, digits=6):
    return float("%.*g" % (digits, x))


def test_plugin_library_exports_the_openmm_entry_points():
    assert os.path.exists(PLUGIN), "run __graft_entry__.build()"
    L = ctypes.CDLL(PLUGIN)
    for sym in ("registerPlatforms", "registerKernelFactories", "registerAGBNPCudaKernelFactories"):
        assert hasattr(L, sym)


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_cuda_available(), reason="only meaningful on a box without a GPU")
def test_cpp_host_fails_loudly_without_a_gpu():
    r = subprocess.run([EXE, "1"], input=_gaussvol_dat(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 1
    assert "no CUDA device" in r.stdout or "CUDA" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("version", [0, 1])
def test_cpp_host_reproduces_the_golden_files(golden, version):
    r = subprocess.run([EXE, str(version)], input=_gaussvol_dat(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    vals = {}
    energies = []
    for line in r.stdout.splitlines():
        k, _, v = line.partition(":")
        if k == "Energy":
            energies.append(float(v))
        else:
            vals[k] = float(v)
    g = golden["v%d" % version]
    assert _sig(energies[0]) == g["energy"]
    assert _sig(energies[1]) == g["energy_displaced"]
    assert abs(vals["Energy Change"] - g["energy_change"]) <= 2e-3        # difference of two float-path energies of ~1e3
    assert abs(vals["Energy Change from Gradient"] - g["energy_change_from_gradient"]) <= 1e-4 * abs(g["energy_change_from_gradient"]) + 1e-7
    if version == 1:
        assert abs(vals["Energy after charge update"] - energies[1]) > 1.0
