#!/usr/bin/env python
"""Stage-by-stage comparison of the CUDA path with the oracle (run on a GPU box: `python tools/stage_check.py [systems]`).
Prints, per system and version, the parity figures of every intermediate the C-ABI exposes."""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import openmm_agbnp_plugin_b200 as plug  # noqa: E402
from openmm_agbnp_plugin_b200 import systems  # noqa: E402
from oracle import portlib  # noqa: E402


def relrms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def gpu_topology(topo_rows):
    """agbnp_b200 TREE_TOPOLOGY dump -> dict path -> ordered child atoms"""
    m = len(topo_rows)
    paths = [None] * m
    kids = {}
    for w in range(m):
        root, par, atom, rank = (int(x) for x in topo_rows[w])
        ppath = (root,) if par < 0 else paths[par]
        paths[w] = ppath + (atom,)
        kids.setdefault(ppath, []).append((rank, atom))
    return {p: [a for _, a in sorted(v)] for p, v in kids.items()}


def check(name, version, method=0, cutoff=1.0, verbose=True):
    s = systems.load(name) if isinstance(name, str) else name
    label = name if isinstance(name, str) else s.get("name", "system")
    pos = systems.float_rounded(s["pos"])
    args = (s["radius"], s["gamma"], s["alpha"], s["charge"], s["ishydrogen"])
    t0 = time.time()
    o = portlib.OracleKernel(version, *args, nonbonded_method=method, cutoff=cutoff)
    e_ref, f_ref = o.execute(pos)
    t_cpu = time.time() - t0
    force = systems.make_force(s, version, method, cutoff)
    ctx = plug.Context(force)
    ctx.setPositions(pos)
    t0 = time.time()
    e = ctx.calcForcesAndEnergy()
    t_gpu = time.time() - t0
    f = ctx.getForces()
    k = ctx.kernel
    res = dict(system=label, version=version, method=method, n=len(pos), e_ref=e_ref, e_gpu=e,
               e_rel=abs(e - e_ref) / abs(e_ref), f_relrms=relrms(f, f_ref), t_cpu=t_cpu, t_gpu_first=t_gpu)
    sc = k.get("SCALARS")
    res["scalars"] = sc
    res["evol1_rel"] = abs(sc[0] - o.scalar("vol_energy1")) / abs(o.scalar("vol_energy1"))
    res["evol2_rel"] = abs(sc[1] - o.scalar("vol_energy2")) / abs(o.scalar("vol_energy2"))
    res["svS_relrms"] = relrms(k.get("SELF_VOLUME_VDW"), o.get("self_volume"))
    res["svL_relrms"] = relrms(k.get("SELF_VOLUME_LARGE"), o.get("self_volume_large"))
    res["tree_nodes"] = (int(k.get("TREE_SIZE")[0]), len(o.tree()["level"]) - 1 - len(pos))
    tg = gpu_topology(k.get("TREE_TOPOLOGY"))
    tr = {p: v for p, v in portlib.tree_topology(o.tree()).items() if len(p) >= 1}
    res["topology_equal"] = (tg == tr)
    if not res["topology_equal"]:
        diff = [p for p in set(tg) | set(tr) if tg.get(p) != tr.get(p)]
        res["topology_diff"] = len(diff)
        res["topology_example"] = [(p, tg.get(p), tr.get(p)) for p in sorted(diff, key=len)[:3]]
    if version == 1:
        res["egb_rel"] = abs(sc[2] - (o.scalar("gb_self") + o.scalar("gb_pair"))) / abs(o.scalar("gb_self") + o.scalar("gb_pair"))
        res["evdw_rel"] = abs(sc[3] - o.scalar("evdw")) / abs(o.scalar("evdw"))
        res["born_relrms"] = relrms(k.get("BORN_RADIUS"), o.get("born_radius"))
        res["born_maxrel"] = float(np.abs(k.get("BORN_RADIUS") / o.get("born_radius") - 1).max())
        res["Y_relrms"] = relrms(k.get("DERIV_Y"), o.get("Y"))
        res["WU_relrms"] = relrms(k.get("DERIV_WU"), o.get("W") + o.get("U"))
        wc = k.get("WORK_COUNTERS")
        res["counters_gpu"] = wc[:7]
        res["counters_cpu"] = [o.counter(c) for c in ("P_gb", "P_q", "C2", "C3", "M")]
    if method == 1:
        pg = k.get("NEIGHBOR_PAIRS")
        pr = portlib.neighbor_pairs(pos.astype(np.float32), cutoff)
        sg = set(map(tuple, pg.tolist())); sr = set(map(tuple, pr.tolist()))
        res["neighbor_equal"] = (sg == sr)
        res["neighbor_count"] = (len(sg), len(sr))
    if verbose:
        for key, val in res.items():
            print("  %-16s %s" % (key, val))
    ctx.kernel.close()
    return res


if __name__ == "__main__":
    names = sys.argv[1:] or ["gaussvol", "trpcage", "rnaseh"]
    for nm in names:
        s = dict(np.load(os.path.join(ROOT, "tests", "golden", "gaussvol.npz"))) if nm == "gaussvol" else nm
        if nm == "gaussvol":
            s["name"] = "gaussvol"
        for v in (0, 1):
            print("== %s v%d NoCutoff" % (nm, v))
            check(s, v)
        print("== %s v1 Cutoff 1.2" % nm)
        check(s, 1, 1, 1.2)
