"""One AGBNP1 evaluation sharded over the GPUs of one box: one process per GPU, `torch.distributed` (NCCL over NVLink 5 /
NVSwitch) for the exchanges between the phases the C-ABI exposes (include/agbnp_b200.h, "multi-GPU plumbing";
SURVEY.md section 8e).  Nothing here computes energies or forces.

    positions  --broadcast from the rank that owns them-->  every rank
    phase 0    overlap trees of the owned roots           -> all-reduce  partial surface-tension gradients + self-volumes (2 x np float4)
    phase 1    Born-radius pair sums, owned units         -> all-reduce  partial sums           (np floats)
    phase 2    Born radii (all atoms) + owned GB tiles    -> all-reduce  partial GB force + Y   (np float4)
    phase 3    bru/brw + derivative pass, owned units     -> all-reduce  partial force + W+U    (np float4)
    phase 4    tree gamma sweep, owned subtrees           -> all-reduce  partial forces (np float4) + energies (8 doubles)
    finish     scatter forces into the caller's sink, total energy

The evaluator below is written against a small "shard kernel" protocol (phase / buffer / finish) so that the exchange
logic runs unchanged on CPU tensors with the gloo backend in tests/test_sharding.py; `CudaShardKernel` is the real one.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .AGBNPplugin import CalcAGBNPForceKernel, OpenMMException

# exchange after each phase: (buffer name, reduce op)
EXCHANGES = (("SELFVOL",), ("BSUM",), ("YQ",), ("WU",), ("FORCE", "ENERGY"))
N_PHASES = len(EXCHANGES)


class _DevPtr:
    """Zero-copy view of a device buffer owned by libagbnp_b200.so (CUDA array interface v2)."""

    def __init__(self, ptr, nbytes, typestr, itemsize):
        self.__cuda_array_interface__ = dict(shape=(nbytes // itemsize,), typestr=typestr, data=(ptr, False), version=2)


class CudaShardKernel:
    """The C-ABI shard entry points of one handle, with the exchange buffers exposed as torch tensors."""

    _TYPES = dict(SELFVOL=("<f4", 4, torch.float32), YQ=("<f4", 4, torch.float32), WU=("<f4", 4, torch.float32),
                  FORCE=("<f4", 4, torch.float32), ENERGY=("<f8", 8, torch.float64), BSUM=("<f4", 4, torch.float32))

    def __init__(self, force, device, shard_rank, shard_count):
        self.kernel = CalcAGBNPForceKernel(CalcAGBNPForceKernel.Name(), None, device, shard_rank, shard_count)
        self.kernel.initialize(None, force)
        self.device = device
        self._buffers = {}
        self.peer = False

    @property
    def handle(self):
        return self.kernel.handle

    def _check(self, rc):
        if rc != _lib.OK:
            raise OpenMMException(self.kernel._err())

    def phase(self, index, d_posq, stream):
        L = _lib.lib()
        self._check(L.agbnp_b200_shard_phase(self.handle, index, C.c_void_p(d_posq.data_ptr() if d_posq is not None else 0),
                                             C.c_void_p(stream)))

    def buffer(self, name):
        if name not in self._buffers:
            L = _lib.lib()
            ptr, nbytes = C.c_void_p(), C.c_size_t()
            self._check(L.agbnp_b200_shard_buffer(self.handle, _lib.BUF[name], C.byref(ptr), C.byref(nbytes)))
            typestr, itemsize, _ = self._TYPES[name]
            self._buffers[name] = torch.as_tensor(_DevPtr(ptr.value, nbytes.value, typestr, itemsize),
                                                  device=torch.device("cuda", self.device))
        return self._buffers[name]

    def setup_peer_exchange(self, group=None):
        """Replace the NCCL all-reduces by the library's one-shot exchange over NVLink peer memory (include/agbnp_b200.h,
        "peer-memory exchange"): export this shard's mailbox, gather everybody's CUDA IPC handles, import them."""
        L = _lib.lib()
        mine = (C.c_ubyte * 64)()
        ok = L.agbnp_b200_peer_export(self.handle, C.cast(mine, C.c_void_p)) == _lib.OK
        world = dist.get_world_size(group)
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(mine) if ok else b"", group=group)
        if ok and all(len(g) == 64 for g in gathered):
            blob = b"".join(gathered)
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            ok = L.agbnp_b200_peer_import(self.handle, C.cast(buf, C.c_void_p), world) == _lib.OK
        else:
            ok = False
        # every shard must take the same path: fall back to NCCL together if any import failed (no peer access)
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok), group=group)
        self.peer = all(flags)
        return self.peer

    @staticmethod
    def setup_peer_local(kernels):
        """All shards in THIS process (one host thread driving several GPUs, or several shards on one GPU in the tests):
        link the mailboxes directly (agbnp_b200_peer_import_local)."""
        L = _lib.lib()
        arr = (C.c_void_p * len(kernels))(*[k.handle for k in kernels])
        for k in kernels:
            k._check(L.agbnp_b200_peer_import_local(k.handle, arr, len(kernels)))
            k.peer = True

    def broadcast(self, d_posq, owner, stream):
        self._check(_lib.lib().agbnp_b200_peer_broadcast(self.handle, C.c_void_p(d_posq.data_ptr()), owner, C.c_void_p(stream)))

    def evaluate_graph(self, d_posq, owner, stream, d_force, layout, padded_n, d_energy):
        """One asynchronous sharded evaluation in one library call (agbnp_b200_shard_evaluate)."""
        self._check(_lib.lib().agbnp_b200_shard_evaluate(self.handle, C.c_void_p(d_posq.data_ptr()), owner, C.c_void_p(stream),
                                                        C.c_void_p(d_force.data_ptr() if d_force is not None else 0), layout, padded_n,
                                                        C.c_void_p(d_energy.data_ptr() if d_energy is not None else 0)))

    def exchange(self, name, stream):
        self._check(_lib.lib().agbnp_b200_peer_exchange(self.handle, _lib.BUF[name], C.c_void_p(stream)))

    def finish(self, stream, d_force, layout, padded_n, d_energy, want_energy):
        """Returns (rc, energy); rc != 0 means this shard overflowed a capacity (already grown) and the evaluation must
        be re-run by every rank."""
        L = _lib.lib()
        e = C.c_double(0.0)
        rc = L.agbnp_b200_shard_finish(self.handle, C.c_void_p(stream), C.c_void_p(d_force.data_ptr() if d_force is not None else 0),
                                       layout, padded_n, C.c_void_p(d_energy.data_ptr() if d_energy is not None else 0),
                                       C.byref(e) if want_energy else None)
        if rc not in (_lib.OK, _lib.ERR_CAPACITY):
            self._check(rc)
        return rc, e.value

    def close(self):
        self._buffers.clear()
        self.kernel.close()


class ShardedEvaluator:
    """Runs the phases of one evaluation on this rank's shard kernel and the collectives between them."""

    def __init__(self, shard_kernel, group=None, position_owner=0):
        self.k = shard_kernel
        self.group = group
        self.position_owner = position_owner
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.collectives = 0            # issued by this rank (diagnostics / tests)

    def _all_reduce(self, t, op=dist.ReduceOp.SUM):
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)
            self.collectives += 1

    def evaluate(self, posq, stream=0, d_force=None, layout=0, padded_n=0, d_energy=None, want_energy=True,
                 broadcast_positions=True, max_attempts=8):
        """posq: float4-per-atom tensor on this rank's device (contents only matter on `position_owner` when
        broadcast_positions is set).  Returns the total energy (want_energy) or None (asynchronous)."""
        if self.world > 1 and not want_energy and broadcast_positions and getattr(self.k, "peer", False) and hasattr(self.k, "evaluate_graph"):
            # asynchronous evaluation over peer memory: broadcast, phases and exchanges enqueued by one library call
            self.k.evaluate_graph(posq, self.position_owner, stream, d_force, layout, padded_n, d_energy)
            self.collectives += 1 + len(EXCHANGES)       # broadcast + one exchange per phase (the last carries forces, energies and status)
            return None
        if self.world > 1 and broadcast_positions:
            if getattr(self.k, "peer", False):
                self.k.broadcast(posq, self.position_owner, stream)
            else:
                dist.broadcast(posq, src=self.position_owner, group=self.group)
            self.collectives += 1
        for _ in range(max_attempts):
            peer = self.world > 1 and getattr(self.k, "peer", False)
            for ph in range(N_PHASES):
                self.k.phase(ph, posq if ph == 0 else None, stream)
                for name in EXCHANGES[ph]:
                    if peer:
                        self.k.exchange(name, stream)       # push to the peers' mailboxes + sum, on the evaluation's stream
                        self.collectives += 1
                    else:
                        self._all_reduce(self.k.buffer(name))
            rc, e = self.k.finish(stream, d_force, layout, padded_n, d_energy, want_energy)
            if not want_energy:
                return None
            # agree on the outcome: if any shard overflowed a capacity, every rank re-runs
            flag = torch.tensor([float(rc != 0)], dtype=torch.float32, device=posq.device)
            self._all_reduce(flag, dist.ReduceOp.MAX)
            if float(flag.item()) == 0.0:
                return e
        raise OpenMMException("agbnp_b200: capacity growth did not converge on some shard")


def owned_rows(n_blocks, rank, world):
    """Contiguous row-block range a shard owns in the derivative pass (mirror of pair_common() in csrc/agbnp_b200.cu)."""
    per = (n_blocks + world - 1) // world
    begin = min(n_blocks, per * rank)
    return begin, min(n_blocks, begin + per)


def owned_roots(n_heavy, n_heavy_blocks, rank, world, tile=32):
    """Positions in the tree work-item list (most expensive first) that a shard processes: a block-cyclic deal in blocks of
    32 (mirror of k_tree).  With no root split into parts the items are the heavy roots themselves."""
    out = []
    for b in range(rank, n_heavy_blocks, world):
        out.extend(range(b * tile, min((b + 1) * tile, n_heavy)))
    return np.array(out, dtype=np.int64)
