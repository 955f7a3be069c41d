#!/usr/bin/env python
"""Energy / force / tree-size parity of the CUDA path against the COMMITTED outputs of the compiled reference
(tests/golden/ref_outputs*.npz), NoCutoff, AGBNP1: python tools/quick_parity.py [system ...]   (GPU box; variants through AGBNP_B200_LIB)"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import openmm_agbnp_plugin_b200 as plug  # noqa: E402
from openmm_agbnp_plugin_b200 import systems  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
small, large = np.load(os.path.join(G, "ref_outputs.npz")), np.load(os.path.join(G, "ref_outputs_large.npz"))
for nm in sys.argv[1:] or ["trpcage", "rnaseh", "1dwc", "2clr", "hivrt_standin"]:
    gold = small if nm + "_v1_energy" in small.files else large
    s = systems.load(nm)
    pos = systems.float_rounded(s["pos"])
    ctx = plug.Context(systems.make_force(s, 1, 0, 1.0))
    ctx.setPositions(pos)
    e = ctx.calcForcesAndEnergy()
    f = ctx.getForces()
    e_ref, f_ref = float(gold[nm + "_v1_energy"]), gold[nm + "_v1_forces"]
    sc = ctx.kernel.get("SCALARS")
    print("%-14s dE/E %+.2e  force relrms %.2e  nodes %d/%d  E_gb %.6f" % (
        nm, (e - e_ref) / abs(e_ref), np.sqrt(((f - f_ref) ** 2).sum() / (f_ref ** 2).sum()),
        int(ctx.kernel.get("TREE_SIZE")[0]), int(gold[nm + "_v1_tree_size"]) - 1 - len(pos), sc[2]))
    ctx.kernel.close()
