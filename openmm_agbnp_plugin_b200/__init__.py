"""openmm_agbnp_plugin_b200 -- B200-native (sm_100a) AGBNP1 / GaussVol energy+force path.

Drop-in for one path of Gallicchio-Lab/openmm_agbnp_plugin: what CalcAGBNPForceKernel::initialize / execute /
copyParametersToContext do (openmmapi/include/AGBNPKernels.h:19-47).  The product is csrc/ (hand-written CUDA kernels
behind the C-ABI declared in include/agbnp_b200.h, built into lib/libagbnp_b200.so); the Python here only mirrors the
reference's user-facing interface (AGBNPplugin.AGBNPForce, python/AGBNPPlugin.i:47-85) on top of that C-ABI.
There is no CPU fallback: importing works anywhere, evaluating requires the built library and a CUDA device.
"""
from .AGBNPplugin import AGBNPForce, CalcAGBNPForceKernel, Context, OpenMMException  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["AGBNPForce", "CalcAGBNPForceKernel", "Context", "OpenMMException"]
