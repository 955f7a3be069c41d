#!/usr/bin/env python
"""Tail of the persistent kernels: mean warp end time against the kernel span, from a -DTAIL_DEBUG build:
   tools/build_variant.sh taildbg -DTAIL_DEBUG;  AGBNP_B200_LIB=variants/taildbg/libagbnp_b200.so python tools/tail_probe.py [system ...]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import openmm_agbnp_plugin_b200 as plug
from openmm_agbnp_plugin_b200 import systems, _lib
L = _lib.lib()
names = ["k_born", "k_gb", "k_deriv", "k_tree", "k_tree_gamma"]
for nm in sys.argv[1:] or ["hivrt", "2clr"]:
    s = systems.load(nm); pos = systems.float_rounded(s["pos"])
    ctx = plug.Context(systems.make_force(s, 1, 0, 1.0)); ctx.setPositions(pos)
    for _ in range(6): ctx.calcForcesAndEnergy()
    agg = np.zeros((5, 3))
    R = 10
    for _ in range(R):
        L.agbnp_b200_debug_tail_reset()
        ctx.calcForcesAndEnergy()
        out = (C.c_ulonglong * 128)()
        L.agbnp_b200_debug_tail_read(out)
        a = np.array(list(out), dtype=np.float64).reshape(16, 8)
        for k in range(5):
            t0, t1, ssum, n = a[k, 0], a[k, 1], a[k, 2], a[k, 3]
            agg[k] += [(t1 - t0) / 1e3, (ssum / n - t0) / 1e3, n]
    agg /= R
    print(nm)
    for k in range(5):
        print("  %-13s span %.1f us  mean warp end %.1f us  (%.0f %% of the span)  warps %d" % (names[k], agg[k, 0], agg[k, 1], 100 * agg[k, 1] / agg[k, 0], agg[k, 2]))
    ctx.kernel.close()
