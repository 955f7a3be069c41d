#!/usr/bin/env python
"""Diagnostic: two shard handles on one GPU through the peer-memory exchange, with host timing per call."""
import ctypes as C, os, sys, time
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import openmm_agbnp_plugin_b200 as plug
from openmm_agbnp_plugin_b200 import systems, _lib, sharding
s = systems.load("1li2"); pos = systems.float_rounded(s["pos"]); n = len(pos)
force = systems.make_force(s, 1, 0, 1.0)
ks = [sharding.CudaShardKernel(force, 0, r, 2) for r in range(2)]
sharding.CudaShardKernel.setup_peer_local(ks)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]; sp = [st.cuda_stream for st in streams]
posq = torch.zeros((n, 4), dtype=torch.float32); posq[:, :3] = torch.from_numpy(pos.astype(np.float32))
d_posq = [posq.cuda(), torch.zeros((n, 4), dtype=torch.float32, device="cuda")]
torch.cuda.synchronize()
L = _lib.lib()
def stats(k):
    st = np.zeros(8); L.agbnp_b200_get(k.handle, _lib.GET["STATS"], st.ctypes.data_as(C.c_void_p), st.nbytes); return st
t0 = time.time()
for r in (0, 1): ks[r].broadcast(d_posq[r], 0, sp[r])
for ph in range(sharding.N_PHASES):
    for r in (0, 1): ks[r].phase(ph, d_posq[r] if ph == 0 else None, sp[r])
    for name in sharding.EXCHANGES[ph]:
        for r in (0, 1): ks[r].exchange(name, sp[r])
for r in (0, 1):
    d_f = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
    rc, e = ks[r].finish(sp[r], d_f, 0, n, None, True)
    print("sync finish shard", r, rc, e, "%.3f s" % (time.time()-t0), flush=True)
for r in (0, 1):
    ps = np.zeros(64); L.agbnp_b200_get(ks[r].handle, _lib.GET["PEER_STATE"], ps.ctypes.data_as(C.c_void_p), ps.nbytes)
    print("after (a): shard", r, "epochs", ps[:7].astype(int), "flags", ps[7:63].reshape(7, 8)[:, :2].astype(int).tolist())
d_e = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in (0, 1)]
d_f = [torch.zeros((n, 3), dtype=torch.float32, device="cuda") for _ in (0, 1)]
torch.cuda.synchronize()
for it in range(6):
    for r in (0, 1):
        t1 = time.time()
        rc = L.agbnp_b200_shard_evaluate(ks[r].handle, C.c_void_p(d_posq[r].data_ptr()), 0, C.c_void_p(sp[r]), C.c_void_p(d_f[r].data_ptr()), 0, n, C.c_void_p(d_e[r].data_ptr()))
        print("eval", it, "shard", r, "rc", rc, "%.3f s" % (time.time()-t1), L.agbnp_b200_last_error(ks[r].handle)[:120] if rc else "", flush=True)
for r in (0, 1):
    t1 = time.time(); rc = L.agbnp_b200_synchronize(ks[r].handle, C.c_void_p(sp[r])); print("sync", r, rc, "%.3f s" % (time.time()-t1), stats(ks[r]))
print(d_e[0].item()/6, d_e[1].item()/6)
for r in (0, 1):
    ps = np.zeros(64); L.agbnp_b200_get(ks[r].handle, _lib.GET["PEER_STATE"], ps.ctypes.data_as(C.c_void_p), ps.nbytes)
    print("shard", r, "epochs", ps[:7].astype(int), "fault", int(ps[63]))
    print("   flags [kind][src 0,1]:", ps[7:63].reshape(7, 8)[:, :2].astype(int).tolist())
