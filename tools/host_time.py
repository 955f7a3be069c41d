#!/usr/bin/env python
"""Host-path (agbnp_b200_execute_host) wall time per evaluation and the library's own breakdown (stderr, every 100 calls):
AGBNP_B200_HOST_TIMING=1 python tools/host_time.py [system] [calls]"""
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import openmm_agbnp_plugin_b200 as plug  # noqa: E402
from openmm_agbnp_plugin_b200 import systems  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "hivrt"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 300
s = systems.load(name)
pos = systems.float_rounded(s["pos"])
ctx = plug.Context(systems.make_force(s, 1, 0, 1.0))
ctx.setPositions(pos)
for _ in range(10):
    ctx.calcForcesAndEnergy()
t0 = time.perf_counter()
for _ in range(calls):
    e = ctx.calcForcesAndEnergy()
t1 = time.perf_counter()
print("%s N=%d: %.1f us per host evaluation (E=%.4f)" % (name, len(pos), (t1 - t0) / calls * 1e6, e))
