#!/usr/bin/env python
"""N-GPU check of one sharded evaluation (run under torchrun): result against the unsharded evaluation on the same GPU, and
the time per evaluation with NCCL all-reduces vs the peer-memory exchange.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py [system]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import openmm_agbnp_plugin_b200 as plug  # noqa: E402
from openmm_agbnp_plugin_b200 import systems, sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "hivrt"
s = systems.hivrt() if name == "hivrt" else systems.load(name)
pos = systems.float_rounded(s["pos"])
n = len(pos)
force = systems.make_force(s, 1, 0, 1.0)
ctx = plug.Context(force, device=local)
ctx.setPositions(pos)
e_ref = ctx.calcForcesAndEnergy()
f_ref = ctx.getForces().copy()
posq = torch.zeros((n, 4), dtype=torch.float32)
posq[:, :3] = torch.from_numpy(pos.astype(np.float32))
d_posq = posq.to(dev)
st = torch.cuda.current_stream().cuda_stream
for mode in ("nccl", "peer"):
    sk = sharding.CudaShardKernel(force, local, rank, world)
    if mode == "peer":
        sk.setup_peer_exchange()
    ev = sharding.ShardedEvaluator(sk, position_owner=0)
    d_f = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    e = ev.evaluate(d_posq, st, d_f, 0, n, None, True)
    f = d_f.cpu().numpy().astype(np.float64)
    erel = abs(e - e_ref) / abs(e_ref)
    frel = float(np.sqrt(((f - f_ref) ** 2).sum() / (f_ref ** 2).sum()))
    for _ in range(10):
        ev.evaluate(d_posq, st, d_f, 0, n, None, False)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    K = 100
    for _ in range(K):
        ev.evaluate(d_posq, st, d_f, 0, n, None, False)
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / K
    print("rank %d/%d %s: E=%.4f (ref %.4f, rel %.2e) force relrms %.2e  %.1f us/eval" % (rank, world, mode, e, e_ref, erel, frel, dt * 1e6), flush=True)
    assert erel < 5e-6 and frel < 1e-5
    sk.close()
dist.barrier()
dist.destroy_process_group()
