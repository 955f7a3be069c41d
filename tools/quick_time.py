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
REP = 50
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
        frc = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream()
        h = ctx.kernel.handle

        def run(k):
            for _ in range(k):
                rc = L.agbnp_b200_execute_device(h, d.data_ptr(), st.cuda_stream, frc.data_ptr(), 0, n, None, None)
                assert rc == 0, L.agbnp_b200_last_error(h)
        run(5)
        assert L.agbnp_b200_synchronize(h, st.cuda_stream) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); run(REP); e1.record(st)
        assert L.agbnp_b200_synchronize(h, st.cuda_stream) == 0
        ms = e0.elapsed_time(e1) / REP
        L.agbnp_b200_profile(h, 0xffffffff)
        run(20)
        sums = (C.c_double * 16)(); cnt = (C.c_int * 16)(); names_p = C.c_char_p()
        nk = L.agbnp_b200_profile_read(h, sums, cnt, 16, C.byref(names_p))
        L.agbnp_b200_profile(h, 0)
        kn = names_p.value.decode().split("\n")
        t2 = time.time()
        for _ in range(20):
            ctx.calcForcesAndEnergy()
        t3 = time.time()
        sc = ctx.kernel.get("SCALARS")
        wc = ctx.kernel.get("WORK_COUNTERS")
        print("%s N=%d method=%d E=%.4f first=%.1fms device=%.3f ms/eval host_e2e=%.3f ms/eval nodes=%d" %
              (nm, n, method, e, (t1 - t0) * 1e3, ms, (t3 - t2) / 20 * 1e3, sc[7]))
        print("   " + "  ".join("%s=%.1fus" % (kn[i], sums[i] / max(cnt[i], 1) * 1e3) for i in range(nk)))
        print("   counters: P_gb=%.4g P_q=%.4g C2=%.4g C3=%.4g M=%.4g tiles_gb=%.4g tiles_q=%.4g" % tuple(wc[:7]))
        ctx.kernel.close()
