#!/usr/bin/env python
"""Energy-term and per-atom by-product comparison of the CUDA path against the oracle (GPU box): tools/parity_breakdown.py [system]"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import openmm_agbnp_plugin_b200 as plug  # noqa: E402
from openmm_agbnp_plugin_b200 import systems  # noqa: E402
from oracle import portlib  # noqa: E402

for nm in sys.argv[1:] or ["2clr"]:
    s = systems.load(nm)
    pos = systems.float_rounded(s["pos"])
    ctx = plug.Context(systems.make_force(s, 1))
    ctx.setPositions(pos)
    e = ctx.calcForcesAndEnergy()
    f = ctx.getForces().copy()
    o = portlib.OracleKernel(1, s["radius"], s["gamma"], s["alpha"], s["charge"], s["ishydrogen"])
    e_ref, f_ref = o.execute(pos)
    sc = ctx.kernel.get("SCALARS")
    ref = dict(vol1=o.scalar("vol_energy1"), vol2=o.scalar("vol_energy2"), gb=o.scalar("gb_self") + o.scalar("gb_pair"), vdw=o.scalar("evdw"))
    got = dict(vol1=sc[0], vol2=sc[1], gb=sc[2], vdw=sc[3])
    print("%s: E=%.6f ref=%.6f rel=%.2e  force relrms=%.2e" % (nm, e, e_ref, abs(e - e_ref) / abs(e_ref),
                                                               np.sqrt(((f - f_ref) ** 2).sum() / (f_ref ** 2).sum())))
    for k in ref:
        print("   %-5s gpu=%.6f ref=%.6f diff=%+.3e (%.1e of |E|)" % (k, got[k], ref[k], got[k] - ref[k], abs(got[k] - ref[k]) / abs(e_ref)))
    print("   gb_self ref=%.6f gb_pair ref=%.6f" % (o.scalar("gb_self"), o.scalar("gb_pair")))
    b, b_ref = ctx.kernel.get("BORN_RADIUS"), o.get("born_radius")
    print("   born radius: max rel %.2e  mean rel (signed) %+.2e" % (np.abs(b / b_ref - 1).max(), (b / b_ref - 1).mean()))
    sv, sv_ref = ctx.kernel.get("SELF_VOLUME_VDW"), o.get("self_volume")
    hv = sv_ref > 0
    print("   self volume: max rel %.2e  mean rel (signed) %+.2e" % (np.abs(sv[hv] / sv_ref[hv] - 1).max(), (sv[hv] / sv_ref[hv] - 1).mean()))
    y, y_ref = ctx.kernel.get("DERIV_Y"), o.get("Y")
    print("   Y: relrms %.2e" % np.sqrt(((y - y_ref) ** 2).sum() / (y_ref ** 2).sum()))
    # where does the GB difference sit?  self term from the GPU's own Born radii, evaluated in double on the host
    kd = 4.184 * 332.0 / 10.0 * (-0.5) * (1.0 - 1.0 / 80.0)
    q = s["charge"]
    self_gpu_B = kd * float((q * q / b).sum())
    self_ref_B = kd * float((q * q / b_ref).sum())
    print("   gb_self from GPU Born radii (double) = %.6f, from oracle radii = %.6f, diff %+.3e" % (self_gpu_B, self_ref_B, self_gpu_B - self_ref_B))
    print("   => gb_pair gpu (total - self(gpuB)) = %.6f vs ref %.6f diff %+.3e" % (got["gb"] - self_gpu_B, o.scalar("gb_pair"), got["gb"] - self_gpu_B - o.scalar("gb_pair")))
    # pair energy recomputed in double on the host from the GPU's Born radii (sample of rows to bound the cost)
    pos64 = pos.astype(np.float64)
    n = len(q)
    rows = np.arange(0, n, max(1, n // 400))
    ep_gpuB = ep_refB = 0.0
    for i in rows:
        d2 = ((pos64 - pos64[i]) ** 2).sum(axis=1)
        m = np.ones(n, dtype=bool); m[i] = False
        for bb, tag in ((b, "g"), (b_ref, "r")):
            t = d2[m] + bb[i] * bb[m] * np.exp(-d2[m] / (4.0 * bb[i] * bb[m]))
            v = kd * float((q[i] * q[m] / np.sqrt(t)).sum())
            if tag == "g":
                ep_gpuB += v
            else:
                ep_refB += v
    print("   sampled pair rows (%d): with GPU radii %.6f, with oracle radii %.6f, diff %+.3e (scaled to all rows: %+.3e)" %
          (len(rows), ep_gpuB, ep_refB, ep_gpuB - ep_refB, (ep_gpuB - ep_refB) * n / len(rows)))
