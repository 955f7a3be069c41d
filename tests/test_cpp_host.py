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
@pytest.mark.parametrize("version,precision", [(0, "single"), (1, "single"), (1, "mixed"), (1, "double")])
def test_cpp_host_reproduces_the_golden_files(golden, version, precision):
    """AGBNPForce -> AGBNPForceImpl -> CudaCalcAGBNPForceKernel on the device buffers of a CudaContext (stand-in with the
    real one's conventions): posq float4 / double4 in the platform's atom order, 64-bit fixed-point force buffer, energy
    buffer in the platform's precision.  The context reorders its atoms twice; the listener must keep every particle's
    force where it belongs."""
    r = subprocess.run([EXE, str(version), precision], input=_gaussvol_dat(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    vals = {}
    energies = []
    reorder = {"Reorder relative energy change": [], "Reorder relative force change": []}
    for line in r.stdout.splitlines():
        k, _, v = line.partition(":")
        if k == "Energy":
            energies.append(float(v))
        elif k in reorder:
            reorder[k].append(float(v))
        else:
            vals[k] = float(v)
    g = golden["v%d" % version]
    # the single-precision energy buffer holds the total as a float: 6 significant digits still agree to one unit
    tol = 1.5e-6 if precision == "single" else 1e-7
    assert abs(energies[0] - g["energy"]) <= 0.51e-5 * abs(g["energy"]) + tol * abs(g["energy"])
    assert abs(energies[1] - g["energy_displaced"]) <= 0.51e-5 * abs(g["energy_displaced"]) + tol * abs(g["energy_displaced"])
    assert abs(vals["Energy Change"] - g["energy_change"]) <= 3e-3        # difference of two float-path energies of ~1e3
    assert abs(vals["Energy Change from Gradient"] - g["energy_change_from_gradient"]) <= 1e-4 * abs(g["energy_change_from_gradient"]) + 1e-7
    assert len(reorder["Reorder relative force change"]) == 2
    assert max(reorder["Reorder relative force change"]) <= 5e-6          # max norm; float summation order is the only difference (measured 1e-6)
    assert max(reorder["Reorder relative energy change"]) <= 2e-6
    if version == 1:
        assert abs(vals["Energy after charge update"] - energies[1]) > 1.0


def test_swig_interface_and_python_mirror_expose_the_same_surface():
    """python/AGBNPplugin.i (the module built with SWIG against OpenMM) and AGBNPplugin.py (the mirror used where OpenMM is
    absent) must offer the same AGBNPForce: method names and the NonbondedMethod enum of the reference's python/AGBNPPlugin.i."""
    import re
    from openmm_agbnp_plugin_b200 import AGBNPplugin
    text = open(os.path.join(PKG, "python", "AGBNPplugin.i")).read()
    assert re.search(r"^%module AGBNPplugin\s*$", text, re.M)
    body = text[text.index("class AGBNPForce"):]
    methods = set(re.findall(r"\b(\w+)\s*\([^;{]*\)\s*(?:const)?\s*;", body)) - {"AGBNPForce"}
    assert {"getNumParticles", "addParticle", "setParticleParameters", "updateParametersInContext", "setCutoffDistance",
            "setNonbondedMethod", "setVersion", "getParticleParameters"} <= methods       # reference python/AGBNPPlugin.i:47-85
    for m in methods:
        assert hasattr(AGBNPplugin.AGBNPForce, m), m
    for name, val in (("NoCutoff", 0), ("CutoffNonPeriodic", 1), ("CutoffPeriodic", 2)):
        assert re.search(r"\b%s\s*=\s*%d\b" % (name, val), body)
        assert getattr(AGBNPplugin.AGBNPForce, name) == val


def test_cmake_project_configures_against_the_stand_in(tmp_path):
    """openmm_agbnp_plugin_b200/CMakeLists.txt (targets agbnp_b200, AGBNPPlugin, AGBNPPluginCUDA as in the reference's
    CMakeLists.txt:93-101 with the CUDA platform in place of OpenCL): configure in stand-in mode; without OpenMM and without
    -DAGBNP_B200_STANDALONE=ON it must stop with a clear message instead of producing a half-configured tree."""
    import shutil
    cmake = shutil.which("cmake")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not cmake or not os.path.exists(nvcc):
        pytest.skip("cmake / nvcc not available")
    r = subprocess.run([cmake, "-S", PKG, "-B", str(tmp_path / "b"), "-DAGBNP_B200_STANDALONE=ON", "-DCMAKE_CUDA_COMPILER=" + nvcc],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    r = subprocess.run([cmake, "-S", PKG, "-B", str(tmp_path / "c"), "-DOPENMM_DIR=/nonexistent", "-DCMAKE_CUDA_COMPILER=" + nvcc],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode != 0 and "OpenMM with its CUDA platform was not found" in r.stdout
