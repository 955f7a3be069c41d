import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import openmm_agbnp_plugin_b200 as plug
from openmm_agbnp_plugin_b200 import systems
from oracle import portlib
from test_gpu_parity import _spaced_cloud, _gpu
from conftest import sys_args
n = 200
pos = systems.float_rounded(_spaced_cloud(n, 1.25, 0.12, 11))
rng = np.random.default_rng(1)
s = dict(radius=np.full(n, 0.17), gamma=np.full(n, 48.9528), alpha=np.full(n, -0.2), charge=rng.normal(0, 0.2, n),
         ishydrogen=np.zeros(n, dtype=np.int32), pos=pos)
o = portlib.OracleKernel(1, *sys_args(s))
e_ref, f_ref = o.execute(pos)
t = o.tree()
ctx, e, f = _gpu(s, pos, 1)
rows = ctx.kernel.get("TREE_TOPOLOGY")
print("nodes", len(rows), "ref", len(t["level"]) - 1 - n, "E", e, e_ref)
bad = [(w, tuple(int(x) for x in rows[w])) for w in range(len(rows)) if rows[w][1] >= w]
print("bad parents", len(bad), bad[:10])
roots = {}
for w in range(len(rows)):
    roots.setdefault(int(rows[w][0]), []).append(w)
for r, ws in list(roots.items())[:3]:
    print("root", r, "n", len(ws), "first", ws[0])
if bad:
    w0 = bad[0][0]
    r0 = int(rows[w0][0])
    print("root of first bad", r0, "size", len(roots[r0]), "offset in root", w0 - roots[r0][0])
print(rows[:24].tolist())
print(rows[14249-6:14249+6].tolist())
lv = t["level"]; par = t["parent"]; at = t["atom"]
# oracle: children of atom slot 1 (root 0)
print("oracle root0 children:", [int(at[c]) for c in range(t["child_start"][1], t["child_start"][1]+t["child_count"][1])][:30])
