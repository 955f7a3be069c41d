"""CPU tests of the host-side mirror of the reference's public surface (AGBNPplugin.AGBNPForce; reference
openmmapi/include/AGBNPForce.h:39-155, openmmapi/src/AGBNPForce.cpp:15-78, python/AGBNPPlugin.i:47-85)."""
import numpy as np
import pytest

import openmm_agbnp_plugin_b200 as plug
from openmm_agbnp_plugin_b200 import systems


def test_defaults():
    f = plug.AGBNPForce()
    assert f.getNumParticles() == 0
    assert f.getNonbondedMethod() == plug.AGBNPForce.NoCutoff == 0
    assert plug.AGBNPForce.CutoffNonPeriodic == 1 and plug.AGBNPForce.CutoffPeriodic == 2
    assert f.getCutoffDistance() == 1.0
    assert f.getVersion() == 1
    assert abs(f.getSolventRadius() - 0.1) < 1e-8


def test_particles_roundtrip():
    f = plug.AGBNPForce()
    assert f.addParticle(0.165, 48.9, -0.3, 0.25, False) == 0
    assert f.addParticle(0.121, 0.0, 0.0, 0.06, True) == 1
    assert f.getNumParticles() == 2
    assert f.getParticleParameters(1) == (0.121, 0.0, 0.0, 0.06, True)
    f.setParticleParameters(0, 0.17, 48.9, -0.2, -0.5, False)
    assert f.getParticleParameters(0) == (0.17, 48.9, -0.2, -0.5, False)
    with pytest.raises(plug.OpenMMException):
        f.getParticleParameters(2)
    with pytest.raises(plug.OpenMMException):
        f.setParticleParameters(-1, 0.1, 0, 0, 0, False)


def test_set_version_range():
    f = plug.AGBNPForce()
    for v in (0, 1, 2):
        f.setVersion(v)
        assert f.getVersion() == v
    for v in (-1, 3):
        with pytest.raises(plug.OpenMMException, match=r"AGBNPForce::setVersion\(\): illegal version number"):
            f.setVersion(v)


def test_method_and_cutoff_setters():
    f = plug.AGBNPForce()
    f.setNonbondedMethod(plug.AGBNPForce.CutoffNonPeriodic)
    f.setCutoffDistance(1.2)
    assert f.getNonbondedMethod() == 1 and f.getCutoffDistance() == 1.2


def test_kernel_name():
    assert plug.CalcAGBNPForceKernel.Name() == "CalcAGBNPForce"      # AGBNPKernels.h:21-23


def test_fixtures_and_standin():
    s = systems.load("2clr")
    assert len(s["radius"]) == 5983 and int((s["ishydrogen"] == 0).sum()) == 3084
    h = systems.hivrt()
    if h["name"].startswith("hivrt-standin"):
        assert len(h["radius"]) == 17949
        # copies in near contact but not overlapping
        assert np.allclose(h["pos"][5983] - h["pos"][0], [5.5, 0, 0])
    j = systems.jitter(s["pos"], 3)
    assert np.abs(j - s["pos"]).max() <= 0.001 and not np.array_equal(j, s["pos"])
    assert np.array_equal(systems.jitter(s["pos"], 3), j)
