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


def _shard_pair_on_one_gpu(force, monkeypatch=None, caps_for_shard1=None):
    """Two shard handles in this process, both on GPU 0, wired through the library's peer-memory exchange."""
    ks = []
    for r in range(2):
        if r == 1 and caps_for_shard1:
            monkeypatch.setenv("AGBNP_B200_INIT_CAPS", caps_for_shard1)
        ks.append(sharding.CudaShardKernel(force, 0, r, 2))
        if r == 1 and caps_for_shard1:
            monkeypatch.delenv("AGBNP_B200_INIT_CAPS")
    sharding.CudaShardKernel.setup_peer_local(ks)
    return ks


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("method,cutoff", [(0, 1.0), (1, 1.2)])
def test_peer_memory_exchange_two_shards_one_gpu(method, cutoff):
    """The multi-GPU path that bench.py --gpus N actually runs -- k_peer_broadcast, k_peer_allreduce and
    agbnp_b200_shard_evaluate -- driven on ONE GPU: two shard handles on two streams exchange through each other's
    mailboxes (linked with agbnp_b200_peer_import_local; between processes the same mailboxes are opened through CUDA IPC).
    Every shard must end with the unsharded result."""
    import openmm_agbnp_plugin_b200 as plug
    from openmm_agbnp_plugin_b200 import systems, _lib
    from conftest import load_system, relrms
    import ctypes as C
    s = load_system("1li2")
    pos = systems.float_rounded(s["pos"])
    n = len(pos)
    force = systems.make_force(s, 1, method, cutoff)
    ctx = plug.Context(force)
    ctx.setPositions(pos)
    e_ref = ctx.calcForcesAndEnergy()
    f_ref = ctx.getForces().copy()
    ks = _shard_pair_on_one_gpu(force)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    sp = [st.cuda_stream for st in streams]
    posq = torch.zeros((n, 4), dtype=torch.float32)
    posq[:, :3] = torch.from_numpy(pos.astype(np.float32))
    d_posq = [posq.cuda(), torch.zeros((n, 4), dtype=torch.float32, device="cuda")]      # only the owner (shard 0) has positions
    torch.cuda.synchronize()

    # (a) phase by phase with explicit exchanges, synchronous finish: settles the capacities
    for attempt in range(4):
        for r in (0, 1):
            ks[r].broadcast(d_posq[r], 0, sp[r])
        for ph in range(sharding.N_PHASES):
            for r in (0, 1):
                ks[r].phase(ph, d_posq[r] if ph == 0 else None, sp[r])
            for name in sharding.EXCHANGES[ph]:
                for r in (0, 1):
                    ks[r].exchange(name, sp[r])
        outs = []
        for r in (0, 1):
            d_f = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
            rc, e = ks[r].finish(sp[r], d_f, 0, n, None, True)
            outs.append((rc, e, d_f))
        if all(rc == 0 for rc, _, _ in outs):
            break
        assert all(rc != 0 for rc, _, _ in outs)          # a fault is seen by BOTH shards
    assert torch.equal(d_posq[0], d_posq[1])              # the broadcast delivered the owner's positions
    for rc, e, d_f in outs:
        assert rc == 0
        assert abs(e - e_ref) <= 2e-6*abs(e_ref)
        assert relrms(d_f.cpu().numpy().astype(np.float64), f_ref) <= 1e-5

    # (b) the one-call asynchronous evaluation, several in flight
    L = _lib.lib()
    d_e = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in (0, 1)]
    d_f = [torch.zeros((n, 3), dtype=torch.float32, device="cuda") for _ in (0, 1)]
    torch.cuda.synchronize()
    nev = 6
    for it in range(nev):
        for r in (0, 1):                                  # the position owner first (single host thread)
            ks[r].evaluate_graph(d_posq[r], 0, sp[r], d_f[r], 0, n, d_e[r])
    for r in (0, 1):
        assert L.agbnp_b200_synchronize(ks[r].handle, C.c_void_p(sp[r])) == 0
    for r in (0, 1):
        assert abs(d_e[r].item()/nev - e_ref) <= 2e-6*abs(e_ref)
        assert relrms(d_f[r].cpu().numpy().astype(np.float64)/nev, f_ref) <= 1e-5
    for k in ks:
        k.close()


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
def test_overflow_on_one_shard_withholds_delivery_on_all(monkeypatch):
    """A capacity overflow on ONE shard of an asynchronous sharded evaluation: no shard may deliver forces that lack the
    other's partial sums, every shard must report the fault for the same evaluation, nobody may hang, and the shards must
    stay in step (the evaluations after the growth are complete and correct)."""
    import openmm_agbnp_plugin_b200 as plug
    from openmm_agbnp_plugin_b200 import systems, _lib
    from conftest import load_system, relrms
    import ctypes as C
    s = load_system("1li2")
    pos = systems.float_rounded(s["pos"])
    n = len(pos)
    force = systems.make_force(s, 1, 0, 1.0)
    ctx = plug.Context(force)
    ctx.setPositions(pos)
    e_ref = ctx.calcForcesAndEnergy()
    f_ref = ctx.getForces().copy()
    ks = _shard_pair_on_one_gpu(force, monkeypatch, caps_for_shard1="64,32,16")     # shard 1 cannot hold its subtrees
    L = _lib.lib()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    sp = [st.cuda_stream for st in streams]
    posq = torch.zeros((n, 4), dtype=torch.float32)
    posq[:, :3] = torch.from_numpy(pos.astype(np.float32))
    d_posq = [posq.cuda(), torch.zeros((n, 4), dtype=torch.float32, device="cuda")]
    torch.cuda.synchronize()

    def run(count):
        d_f = [torch.zeros((n, 3), dtype=torch.float32, device="cuda") for _ in (0, 1)]
        d_e = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in (0, 1)]
        torch.cuda.synchronize()
        rcs = [[], []]
        for it in range(count):
            for r in (0, 1):
                rcs[r].append(L.agbnp_b200_shard_evaluate(ks[r].handle, C.c_void_p(d_posq[r].data_ptr()), 0, C.c_void_p(sp[r]),
                                                          C.c_void_p(d_f[r].data_ptr()), 0, n, C.c_void_p(d_e[r].data_ptr())))
        for r in (0, 1):
            rcs[r].append(L.agbnp_b200_synchronize(ks[r].handle, C.c_void_p(sp[r])))
        return rcs, d_f, d_e

    rcs, d_f, d_e = run(1)
    assert rcs[0] == rcs[1] == [0, _lib.ERR_CAPACITY]        # reported by BOTH shards, by the same call
    for r in (0, 1):
        assert float(d_f[r].abs().max().item()) == 0.0 and d_e[r].item() == 0.0      # nothing was delivered anywhere
        assert b"NOT delivered" in L.agbnp_b200_last_error(ks[r].handle)
    # growth may take more than one round (each fault doubles what overflowed); afterwards everything is delivered
    for attempt in range(8):
        rcs, d_f, d_e = run(5)
        assert rcs[0] == rcs[1]
        if all(rc == 0 for rc in rcs[0]):
            break
    assert all(rc == 0 for rc in rcs[0])
    for r in (0, 1):
        assert abs(d_e[r].item()/5 - e_ref) <= 2e-6*abs(e_ref)
        assert relrms(d_f[r].cpu().numpy().astype(np.float64)/5, f_ref) <= 1e-5
    for k in ks:
        k.close()
