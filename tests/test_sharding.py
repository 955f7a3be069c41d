"""Multi-GPU plumbing.

CPU part (gloo, world size 2): the exchange logic of openmm_agbnp_plugin_b200/sharding.py -- phases, the all-reduce
after each, agreement on a capacity re-run -- against a mock shard kernel whose partial buffers have known totals.
GPU part (one device): two shard handles (rank 0/1 of 2) driven in lock-step with the all-reduce emulated by adding their
exported buffers; the result must equal the unsharded evaluation.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openmm_agbnp_plugin_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class MockShardKernel:
    """Same protocol as sharding.CudaShardKernel on CPU tensors: phase p deposits (rank+1)*(p+1) into its exchange
    buffers; finish reports a capacity overflow on `fail_rank` for the first `fail_times` attempts."""

    SIZES = dict(SELFVOL=16, BSUM=4, YQ=8, WU=8, FORCE=8, ENERGY=8)

    def __init__(self, rank, fail_rank=-1, fail_times=0):
        self.rank, self.fail_rank, self.fail_left = rank, fail_rank, fail_times
        self.bufs = {k: torch.zeros(n, dtype=torch.float64 if k == "ENERGY" else torch.float32) for k, n in self.SIZES.items()}
        self.phases = []
        self.seen = {}

    def phase(self, index, d_posq, stream):
        if index == 0:
            for b in self.bufs.values():
                b.zero_()
            self.posq = d_posq.clone()
        else:
            # what the previous exchange delivered must already be the global total
            for name in sharding.EXCHANGES[index-1]:
                self.seen[name] = self.bufs[name].clone()
        for name in sharding.EXCHANGES[index]:
            self.bufs[name] += float((self.rank+1)*(index+1))
        self.phases.append(index)

    def buffer(self, name):
        return self.bufs[name]

    def finish(self, stream, d_force, layout, padded_n, d_energy, want_energy):
        for name in sharding.EXCHANGES[-1]:
            self.seen[name] = self.bufs[name].clone()
        if self.rank == self.fail_rank and self.fail_left > 0:
            self.fail_left -= 1
            return -3, 0.0
        return 0, float(self.bufs["ENERGY"][0])


def _worker(rank, world, port, fail_rank, fail_times, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = MockShardKernel(rank, fail_rank, fail_times)
        ev = sharding.ShardedEvaluator(k, position_owner=0)
        posq = torch.full((5, 4), float(rank+7))
        e = ev.evaluate(posq, 0, None, 0, 5, None, True)
        tot = sum(r+1 for r in range(world))
        ok = True
        ok &= bool(torch.all(k.posq == 7.0))                                   # positions came from the owner
        for p, names in enumerate(sharding.EXCHANGES):
            for name in names:
                ok &= bool(torch.all(k.seen[name] == tot*(p+1)))               # every exchange delivered the global sum
        ok &= e == tot*len(sharding.EXCHANGES)
        ok &= k.phases == list(range(sharding.N_PHASES))*(fail_times+1)        # everyone re-ran together
        # broadcast + one all-reduce per exchanged buffer + the agreement flag, per attempt
        per_attempt = sum(len(x) for x in sharding.EXCHANGES) + 1
        ok &= ev.collectives == 1 + per_attempt*(fail_times+1)
        out[rank] = int(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fail_times", [0, 2])
def test_exchange_logic_gloo_world2(fail_times):
    world = 2
    port = _free_port()
    out = mp.Array("i", [0]*world)
    procs = [mp.Process(target=_worker, args=(r, world, port, 1, fail_times, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1]*world


def test_ownership_partitions_cover_everything_once():
    for n_blocks, world in [(1, 2), (7, 2), (562, 8), (290, 4)]:
        rows = [sharding.owned_rows(n_blocks, r, world) for r in range(world)]
        covered = np.zeros(n_blocks, dtype=int)
        for b, e in rows:
            covered[b:e] += 1
        assert np.all(covered == 1)
    for nh, world in [(1, 2), (100, 2), (9252, 8)]:
        nhb = (nh+31)//32
        allr = np.concatenate([sharding.owned_roots(nh, nhb, r, world) for r in range(world)])
        assert sorted(allr.tolist()) == list(range(nh))


@pytest.mark.gpu
@pytest.mark.parametrize("method,cutoff", [(0, 1.0), (1, 1.2)])
def test_two_shards_on_one_gpu_equal_the_unsharded_result(method, cutoff):
    import openmm_agbnp_plugin_b200 as plug
    from openmm_agbnp_plugin_b200 import systems
    from conftest import load_system, relrms
    s = load_system("1li2")
    pos = systems.float_rounded(s["pos"])
    n = len(pos)
    force = systems.make_force(s, 1, method, cutoff)
    ctx = plug.Context(force)
    ctx.setPositions(pos)
    e_ref = ctx.calcForcesAndEnergy()
    f_ref = ctx.getForces().copy()

    world = 2
    ks = [sharding.CudaShardKernel(force, 0, r, world) for r in range(world)]
    posq = torch.zeros((n, 4), dtype=torch.float32)
    posq[:, :3] = torch.from_numpy(pos.astype(np.float32))
    d_posq = posq.cuda()
    stream = torch.cuda.current_stream().cuda_stream
    for ph in range(sharding.N_PHASES):
        for k in ks:
            k.phase(ph, d_posq if ph == 0 else None, stream)
        torch.cuda.synchronize()
        for name in sharding.EXCHANGES[ph]:                      # the all-reduce, by hand
            tot = ks[0].buffer(name) + ks[1].buffer(name)
            for k in ks:
                k.buffer(name).copy_(tot)
    outs = []
    for k in ks:
        d_f = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
        rc, e = k.finish(stream, d_f, 0, n, None, True)
        assert rc == 0
        outs.append((e, d_f.cpu().numpy().astype(np.float64)))
    for e, f in outs:                                            # every rank ends with the full result
        assert abs(e - e_ref) <= 2e-6*abs(e_ref)
        assert relrms(f, f_ref) <= 1e-5
    for k in ks:
        k.close()


@pytest.mark.gpu
def test_two_shards_with_tree_reuse(monkeypatch):
    """Opt-in tree reuse under sharding: every shard keeps and rescans the subtrees of the items it owns; the exchanged
    sums must still add up to the unsharded result on the rescanned frames."""
    import openmm_agbnp_plugin_b200 as plug
    from openmm_agbnp_plugin_b200 import systems
    from conftest import load_system, relrms
    s = load_system("1li2")
    pos = systems.float_rounded(s["pos"])
    n = len(pos)
    force = systems.make_force(s, 1, 0, 1.0)
    ctx = plug.Context(force)                                    # rebuilds every evaluation
    monkeypatch.setenv("AGBNP_B200_TREE_REUSE", "4")
    world = 2
    ks = [sharding.CudaShardKernel(force, 0, r, world) for r in range(world)]
    monkeypatch.delenv("AGBNP_B200_TREE_REUSE")
    stream = torch.cuda.current_stream().cuda_stream
    for frame in range(3):                                       # build, rescan, rescan
        if frame:
            pos = systems.float_rounded(systems.jitter(pos, 900 + frame))
        ctx.setPositions(pos)
        e_ref = ctx.calcForcesAndEnergy()
        f_ref = ctx.getForces().copy()
        posq = torch.zeros((n, 4), dtype=torch.float32)
        posq[:, :3] = torch.from_numpy(pos.astype(np.float32))
        d_posq = posq.cuda()
        for ph in range(sharding.N_PHASES):
            for k in ks:
                k.phase(ph, d_posq if ph == 0 else None, stream)
            torch.cuda.synchronize()
            for name in sharding.EXCHANGES[ph]:
                tot = ks[0].buffer(name) + ks[1].buffer(name)
                for k in ks:
                    k.buffer(name).copy_(tot)
        for k in ks:
            d_f = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
            rc, e = k.finish(stream, d_f, 0, n, None, True)
            assert rc == 0
            assert abs(e - e_ref) <= 1e-5*abs(e_ref)
            assert relrms(d_f.cpu().numpy().astype(np.float64), f_ref) <= 1e-4
    for k in ks:
        k.close()
