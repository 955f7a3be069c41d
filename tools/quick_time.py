#!/usr/bin/env python
"""Quick per-kernel timing of the CUDA path (GPU box): python tools/quick_time.py [system ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import openmm_agbnp_plugin_b200 as plug  # noqa: E402
from openmm_agbnp_plugin_b200 import systems, _lib  # noqa: E402

names = sys.argv[1:] or ["trpcage", "rnaseh", "2clr", "hivrt"]
L = _lib.lib()
for nm in names:
    for method, cutoff in ((0, 1.0), (1, 1.2)):
        s = systems.load(nm)
        pos = systems.float_rounded(s["pos"])
        ctx = plug.Context(systems.make_force(s, 1, method, cutoff))
        ctx.setPositions(pos)
        t0 = time.time(); e = ctx.calcForcesAndEnergy(); t1 = time.time()
        n = len(pos)
        posq = torch.zeros((n, 4), dtype=torch.float32)
        posq[:, :3] = torch.from_numpy(pos.astype(np.float32))
        d = posq.cuda()
        ms = C.c_float(0)
        rc = L.agbnp_b200_time_device(ctx.kernel.handle, C.c_void_p(d.data_ptr()), 50, C.byref(ms))
        assert rc == 0, L.agbnp_b200_last_error(ctx.kernel.handle)
        kt = (C.c_float * 16)()
        names_p = C.c_char_p()
        nk = L.agbnp_b200_kernel_times(ctx.kernel.handle, 20, C.c_void_p(d.data_ptr()), kt, 16, C.byref(names_p))
        kn = names_p.value.decode().split("\n")
        t2 = time.time()
        for _ in range(20):
            ctx.calcForcesAndEnergy()
        t3 = time.time()
        sc = ctx.kernel.get("SCALARS")
        print("%s N=%d method=%d E=%.4f first=%.1fms device=%.3f ms/eval host_e2e=%.3f ms/eval nodes=%d" %
              (nm, n, method, e, (t1 - t0) * 1e3, ms.value, (t3 - t2) / 20 * 1e3, sc[7]))
        print("   " + "  ".join("%s=%.1fus" % (kn[i], kt[i] * 1e3) for i in range(nk)))
        ctx.kernel.close()
