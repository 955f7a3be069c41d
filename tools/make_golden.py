#!/usr/bin/env python
"""Build the committed golden fixtures under tests/golden/ (run where /root/reference exists).

  gaussvol.npz     the reference's own test input platforms/reference/tests/gaussvol.dat, converted with the unit and
                   alpha rules of TestReferenceAGBNPForce.cpp:47-70
  golden.json      the numbers printed in platforms/reference/tests/{v0,v1}.reference (the reference's golden outputs)
  ref_outputs.npz  full-precision outputs of the compiled, unmodified reference (oracle/_ref) on gaussvol, trpcage and
                   rnaseh with float-rounded positions: energies, forces, self-volumes, Born radii, tree sizes
  ref_outputs_large.npz  the same for the BASELINE systems 1dwc and 2clr and for the full-size HIV-RT stand-in "2clr x 3"
                   (N = 17 949; systems.hivrt()), AGBNP1 only (ReferenceAGBNPKernels.cpp:274-795): ~15 s of CPU
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import reflib  # noqa: E402

REF_TESTS = "/root/reference/platforms/reference/tests"
OUT = os.path.join(ROOT, "tests", "golden")


def load_gaussvol_dat(path):
    rows = open(path).read().split()
    n = int(rows[0])
    a = np.array(rows[1:1 + 8 * n], dtype=float).reshape(n, 8)
    ang2nm, kcal = 0.1, 4.184
    pos = a[:, 1:4] * ang2nm
    radius = a[:, 4] * ang2nm
    charge = a[:, 5].copy()
    gamma = a[:, 6] * kcal / (ang2nm * ang2nm)
    ish = (a[:, 7] > 0).astype(np.int32)
    # TestReferenceAGBNPForce.cpp:51-68
    sigmaw = 3.15365 * ang2nm
    epsilonw = 0.155 * kcal
    rho = 0.033428 / ang2nm ** 3
    epsilon_lj = 0.155 * kcal
    sij = np.sqrt(sigmaw * 2.0 * radius)
    eij = np.sqrt(epsilonw * epsilon_lj)
    alpha = -16.0 * np.pi * rho * eij * sij ** 6 / 3.0
    return dict(pos=pos, radius=radius, gamma=gamma, alpha=alpha, charge=charge, ishydrogen=ish)


def main():
    os.makedirs(OUT, exist_ok=True)
    g = load_gaussvol_dat(os.path.join(REF_TESTS, "gaussvol.dat"))
    np.savez_compressed(os.path.join(OUT, "gaussvol.npz"), **g)
    # the same input in the reference's own text format (TestReferenceAGBNPForce.cpp:45-58), for the C++ test program
    # (platforms/cuda/tests/TestCudaAGBNPForce < tests/golden/gaussvol.dat; CMake's ctest)
    with open(os.path.join(OUT, "gaussvol.dat"), "w") as fh:
        fh.write("%d\n" % len(g["radius"]))
        for i in range(len(g["radius"])):
            x, y, z = g["pos"][i] * 10.0
            fh.write("%d %.17g %.17g %.17g %.17g %.17g %.17g %d\n" % (i, x, y, z, g["radius"][i] * 10.0, g["charge"][i],
                                                                      g["gamma"][i] * 0.01 / 4.184, int(g["ishydrogen"][i])))

    def nums(path):
        return [float(x) for x in re.findall(r":\s*(-?[0-9.]+(?:e-?[0-9]+)?)", open(path).read())]
    v0 = nums(os.path.join(REF_TESTS, "v0.reference"))
    v1 = nums(os.path.join(REF_TESTS, "v1.reference"))
    # v0.reference: Hsize, E1, E2, SA, Energy, E1', E2', SA', Energy', Change, ChangeFromGradient
    golden = {
        "source": "platforms/reference/tests/v0.reference, v1.reference (input gaussvol.dat; displaced atom 121, +2e-3 nm in y)",
        "v0": {"vol_energy1": v0[1], "vol_energy2": v0[2], "energy": v0[4], "energy_displaced": v0[8],
               "energy_change": v0[9], "energy_change_from_gradient": v0[10]},
        "v1": {"energy": v1[1], "energy_displaced": v1[2], "energy_change": v1[3], "energy_change_from_gradient": v1[4]},
        "displaced_atom": 121, "displaced_axis": 1, "displacement_nm": 2e-3,
    }
    json.dump(golden, open(os.path.join(OUT, "golden.json"), "w"), indent=1)
    print(golden)

    out = {}
    systems = {"gaussvol": g}
    for name in ("trpcage", "rnaseh"):
        s = np.load(os.path.join(OUT, "systems", name + ".npz"))
        systems[name] = {k: s[k] for k in s.files}
    for name, s in systems.items():
        pos = s["pos"].astype(np.float32).astype(np.float64)
        for v in (0, 1):
            k = reflib.ReferenceKernel(v, s["radius"], s["gamma"], s["alpha"], s["charge"], s["ishydrogen"])
            e, f = k.execute(pos)
            out["%s_v%d_energy" % (name, v)] = np.array(e)
            out["%s_v%d_forces" % (name, v)] = f
            out["%s_v%d_tree_size" % (name, v)] = np.array(k.tree_size())
            if v == 1:
                out["%s_self_volume" % name] = k.get("self_volume")
                out["%s_born_radius" % name] = k.get("born_radius")
            print(name, v, "%.10g" % e, k.tree_size())
            k.close()
    np.savez_compressed(os.path.join(OUT, "ref_outputs.npz"), **out)

    from openmm_agbnp_plugin_b200 import systems as psys
    big = {}
    for name in ("1dwc", "2clr", "hivrt_standin"):
        s = psys.load(name)
        pos = psys.float_rounded(s["pos"])
        k = reflib.ReferenceKernel(1, s["radius"], s["gamma"], s["alpha"], s["charge"], s["ishydrogen"])
        e, f = k.execute(pos)
        big["%s_v1_energy" % name] = np.array(e)
        big["%s_v1_forces" % name] = f
        big["%s_v1_tree_size" % name] = np.array(k.tree_size())
        big["%s_self_volume" % name] = k.get("self_volume")
        big["%s_born_radius" % name] = k.get("born_radius")
        print(name, 1, "%.10g" % e, k.tree_size())
        k.close()
    np.savez_compressed(os.path.join(OUT, "ref_outputs_large.npz"), **big)


if __name__ == "__main__":
    main()
