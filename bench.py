#!/usr/bin/env python
"""bench.py -- AGBNP1 energy+force evaluations per second on HIV-RT (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            this repo's sm_100a path (N=1: one GPU; N>1 under torchrun:
                                                              one evaluation sharded over N GPUs, NCCL exchanges)
    python bench.py --impl reference --gpus N ...            the reference's own Reference-platform CPU code (oracle/_ref,
                                                              compiled unmodified from /root/reference) on the host cores

A step is ONE energy+force evaluation of the workload (HIV-RT, AGBNP1 version 1, NoCutoff -- the only method the
Reference platform implements) on freshly jittered coordinates (+-0.001 nm, seeded), so every step rebuilds the overlap
tree as an MD step would.  `value` times the evaluation with positions already resident in HBM (CUDA events on the
launching stream, one bracket per step, L2 flushed between steps outside the brackets); `e2e` times the plugin-level
call with HOST buffers (positions up, forces + energy down, every step).  DESIGN.md section "Measurement" has the details.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_JSON_OUT = sys.stdout
METRIC = "AGBNP1 energy+force evals/s on HIV RT"
UNIT = "evals/s"
NS_PER_DAY_PER_EVAL_PER_S = 0.0864          # dt = 1 fs, one evaluation per step (example/hivrt_benchmark.py:20)
JITTER_SETS = 8
FLUSH_BYTES = 256 << 20                     # > 126 MB L2
TREE_STORE_BYTES_PER_NODE = 34.0            # persisted per overlap-tree node (two float4 + a sibling rank): DESIGN.md 2.1

# algorithmic work per unit (counted from the reference source)
FLOP_GB, MUFU_GB = 42.0, 2.0                # per GB pair
FLOP_Q, MUFU_Q = 86.0, 3.0                  # per directed screening pair (Born pass 28+1, derivative pass 58+2)
FLOP_CAND, MUFU_CAND = 20.0, 3.0            # per overlap candidate examined
FLOP_NODE, MUFU_NODE = 190.0, 5.0           # per tree node: build 25 + rescan 45 + three sweeps 120; 1 + 4 MUFU


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_WORKLOAD = {"name": "hivrt", "method": 0, "cutoff": 1.0}     # BASELINE.json metric: HIV-RT, NoCutoff


def workload():
    from openmm_agbnp_plugin_b200 import systems
    if _WORKLOAD["name"] == "hivrt":
        s = systems.hivrt()
    else:
        s = systems.load(_WORKLOAD["name"])          # secondary configs (BASELINE configs 2-3): committed fixtures
        s["name"] = _WORKLOAD["name"]
    s["pos"] = systems.float_rounded(s["pos"])
    return s


def jittered(pos, k):
    from openmm_agbnp_plugin_b200 import systems
    return systems.float_rounded(systems.jitter(pos, 20261018 + k))


_SAMPLER_SRC = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
bits = [("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
        ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap")]
print("max", mx, flush=True)
while True:
    try:
        c = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        names = ",".join(n for n, a in bits if r & getattr(nv, a, 0))
        print("s", "%.6f" % time.time(), c, names, flush=True)
    except Exception:
        pass
    time.sleep(0.004)
"""


class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs, sampled by a SEPARATE PROCESS (NVML): a
    sampling thread in this process would compete with the launch loop for the interpreter lock, and an NVML call can take
    tens of milliseconds on a busy 8-GPU node.  The samples carry wall-clock stamps; result() keeps those
    that fall inside [t0, t1] (or, if the region was shorter than a sampling period, the nearest ones around it)."""

    def __init__(self, index):
        import subprocess
        self.index = index
        try:
            self.p = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            first = self.p.stdout.readline().split()            # "max <MHz>": the sampler is up
            self.max_mhz = float(first[1]) if len(first) == 2 and first[0] == "max" else None
        except Exception:
            self.p, self.max_mhz = None, None

    def result(self, t0, t1):
        import subprocess
        rows = []
        if self.p is not None:
            time.sleep(0.012)                                   # let a few samples land after the region
            self.p.terminate()
            try:
                out, _ = self.p.communicate(timeout=5)
            except Exception:
                self.p.kill(); out = ""
            for ln in out.splitlines():
                f = ln.split()
                if len(f) >= 3 and f[0] == "s":
                    rows.append((float(f[1]), float(f[2]), f[3].split(",") if len(f) > 3 and f[3] else []))
        inside = [r for r in rows if t0 <= r[0] <= t1]
        note = "sampled by a separate process during the timed region"
        if not inside and rows:
            inside = sorted(rows, key=lambda r: min(abs(r[0]-t0), abs(r[0]-t1)))[:2]
            note = "timed region shorter than the sampling period: the two samples nearest to it"
        if not inside:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     stdout=subprocess.PIPE, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 0, "note": "nvidia-smi sample after the region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({x for r in inside for x in r[2]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(inside), "note": note}


# ------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own Reference-platform code on the host cores
# ------------------------------------------------------------------------------------------------------------------
_W = {}


def _ref_backend():
    from oracle import reflib, portlib
    if reflib.available():
        return "reference", reflib.ReferenceKernel
    return "port", (lambda v, *a: portlib.OracleKernel(v, *a))


def _worker_init(m):
    s = workload()
    kind, make = _ref_backend()
    _W["pos"] = s["pos"][:m]
    _W["k"] = make(1, s["radius"][:m], s["gamma"][:m], s["alpha"][:m], s["charge"][:m], s["ishydrogen"][:m].astype(np.int32))


def _worker_eval(seed):
    t0 = time.perf_counter()
    e, f = _W["k"].execute(jittered(_W["pos"], seed))
    return time.perf_counter() - t0, float(e)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    """bench.py --impl reference: the reference's own Reference-platform code (oracle/_ref: gaussvol.cpp, AGBNPUtils.cpp,
    ReferenceAGBNPKernels.cpp compiled unmodified) on the host cores.  The Reference platform is serial, so "all the host
    threads it can use" is one independent evaluation per core: a step = `cores` full-size evaluations in parallel, each
    on freshly jittered coordinates of the SAME workload as the GPU arm (all N atoms -- about 11 s per evaluation for
    HIV-RT); value = evaluations completed per second.  Only if K+W steps of that would pass ~15 minutes is a prefix of the
    atoms used instead, and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    s = workload()
    n = len(s["pos"])
    kind, make = _ref_backend()
    # load the library and run a small evaluation in THIS process too: the driver records which native libraries the
    # bench process itself loaded (the timed evaluations run in forked workers that inherit it)
    tiny = make(1, s["radius"][:64], s["gamma"][:64], s["alpha"][:64], s["charge"][:64], s["ishydrogen"][:64].astype(np.int32))
    tiny.execute(s["pos"][:64])
    cores = max(1, host_cores())
    steps, warmup = args.steps, args.warmup
    budget_s = 900.0
    t_full = 1.25 * 3.1 * (n / 5983.0) ** 2                # s per full evaluation with every core busy (measured: 2clr 3.1 s alone)
    per_step = budget_s / max(1, steps + warmup)
    m = n if t_full <= per_step else max(1000, min(n, int(n * (per_step / t_full) ** 0.5)))
    scale = (n / m) ** 2
    with ProcessPoolExecutor(cores, mp_context=mp.get_context("fork"), initializer=_worker_init, initargs=(m,)) as pool:
        seeds = iter(range(10 ** 6))
        for _ in range(warmup):
            list(pool.map(_worker_eval, [next(seeds) for _ in range(cores)]))
        t0 = time.perf_counter()
        for _ in range(steps):
            list(pool.map(_worker_eval, [next(seeds) for _ in range(cores)]))
        wall = time.perf_counter() - t0
    ms_per_step = wall / steps * 1e3
    value = cores * steps / (wall * scale)
    sample = ("full workload (all %d atoms), %d evaluations in parallel per step, one per core" % (n, cores) if m == n else
              "first %d of %d atoms per evaluation (one evaluation per core per step); time scaled by (N/m)^2 = %.2f" % (m, n, scale))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic" if s.get("name", "").startswith("hivrt-standin") else s.get("name", "hivrt"),
            "config": config_dict(s, args, "host CPU, Reference platform, %d cores" % cores),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ns_per_day": value * NS_PER_DAY_PER_EVAL_PER_S, "gpu_launches": 0, "same_workload_as_gpu_arm": m == n,
            "evaluations_per_step": cores}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def config_dict(s, args, parallelism):
    meth = "NoCutoff" if _WORKLOAD["method"] == 0 else "CutoffNonPeriodic %.2f nm" % _WORKLOAD["cutoff"]
    return {"workload": "%s AGBNP1 (setVersion 1) energy+force, %s, N=%d atoms" % (s.get("name", "hivrt"), meth, len(s["pos"])),
            "n_atoms": int(len(s["pos"])), "nonbonded_method": meth, "version": 1,
            "inputs": "positions jittered +-0.001 nm per step (seeded, %d sets)" % JITTER_SETS,
            "l2": "256 MiB memset between timed evaluations (outside the per-step brackets), in the device-resident AND the end-to-end loop",
            "parallelism": parallelism,
            **({"tree_reuse_interval": _WORKLOAD["tree_reuse"],
                "tree_reuse_note": "opt-in, NOT the reference's semantics: tree topology kept between builds (SURVEY 8f-3)"}
               if _WORKLOAD.get("tree_reuse") else {})}


def cpu_baseline(s, e_gpu, f_gpu):
    """One full-size Reference-platform evaluation on one host core (bounded: ~30 s); doubles as the full-size parity check."""
    kind, make = _ref_backend()
    k = make(1, s["radius"], s["gamma"], s["alpha"], s["charge"], s["ishydrogen"].astype(np.int32))
    t0 = time.perf_counter()
    e, f = k.execute(s["pos"])
    dt = time.perf_counter() - t0
    f = np.asarray(f).reshape(-1, 3)
    parity = {"energy_rel_err": abs(e_gpu - e) / abs(e), "force_rel_rms": float(np.sqrt(((f_gpu - f) ** 2).sum() / (f ** 2).sum())),
              "energy_gpu": e_gpu, "energy_cpu": float(e), "against": kind, "tolerance": {"energy_rel": 1e-5, "force_rel_rms": 1e-4}}
    base = {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "1 full-size evaluation (N=%d, unjittered positions), %.1f s on one core; the Reference platform is serial" % (len(s["pos"]), dt)}
    return base, parity


# ------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------------
def b200_arm(args):
    import torch
    import torch.distributed as dist
    import openmm_agbnp_plugin_b200 as plug
    from openmm_agbnp_plugin_b200 import systems, _lib, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log("bench: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    s = workload()
    n = len(s["pos"])
    sharded = world > 1 and args.mode == "shard"
    force = systems.make_force(s, 1, _WORKLOAD["method"], _WORKLOAD["cutoff"])

    # device-resident inputs: JITTER_SETS jittered coordinate sets as float4
    posq_sets = []
    for k in range(JITTER_SETS):
        p = torch.zeros((n, 4), dtype=torch.float32)
        p[:, :3] = torch.from_numpy(jittered(s["pos"], k).astype(np.float32))
        posq_sets.append(p.to(dev))
    d_force = torch.zeros((n, 3), dtype=torch.float32, device=dev)
    d_energy = torch.zeros(1, dtype=torch.float64, device=dev)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    if sharded:
        sk = sharding.CudaShardKernel(force, local, rank, world)
        if os.environ.get("AGBNP_B200_PEER", "1") != "0":
            if not sk.setup_peer_exchange():         # one-shot all-reduces over NVLink peer memory instead of NCCL
                log("bench: peer-memory exchange unavailable on this box, using NCCL all-reduces")
        ev = sharding.ShardedEvaluator(sk, position_owner=0)
        handle = sk.handle

        def one_eval(k, want_energy=False):
            try:
                return ev.evaluate(posq_sets[k % JITTER_SETS], sp, d_force, 0, n, d_energy, want_energy)
            except plug.OpenMMException:
                if not tolerant[0]:
                    raise
                faults[0] += 1
    else:
        ctx = plug.Context(force, device=local)
        handle = ctx.kernel.handle

        def one_eval(k, want_energy=False):
            e = C.c_double(0.0)
            rc = L.agbnp_b200_execute_device(handle, posq_sets[k % JITTER_SETS].data_ptr(), sp, d_force.data_ptr(), 0, n,
                                             d_energy.data_ptr(), C.byref(e) if want_energy else None)
            if rc != 0:
                # an asynchronous call reports the overflow of an EARLIER evaluation (its own was enqueued): fine while warming up
                if not (tolerant[0] and rc == _lib.ERR_CAPACITY):
                    raise RuntimeError(L.agbnp_b200_last_error(handle).decode())
                faults[0] += 1
            return e.value

    tolerant, faults = [True], [0]                  # warm-up: deferred overflow reports are expected, the timed region must have none

    def sync():
        rc = L.agbnp_b200_synchronize(handle, sp)
        if rc != 0:
            if not (tolerant[0] and rc == _lib.ERR_CAPACITY):
                raise RuntimeError(L.agbnp_b200_last_error(handle).decode())
            faults[0] += 1
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stats():
        st = np.zeros(8)
        L.agbnp_b200_get(handle, _lib.GET["STATS"], st.ctypes.data_as(C.c_void_p), st.nbytes)
        return st

    # issue-rate peaks of this GPU (roofline denominators), with the SM clock they were measured at
    pk = (C.c_double * 8)()
    L.agbnp_b200_measure_peaks(local, pk, 8)
    peak_fp32 = 2.0 * max(pk[0], pk[1])             # flop/s
    peak_mufu = min(pk[2], pk[3])                   # op/s

    # ---- warm-up, regardless of --warmup: (1) every jitter set once through the synchronous path, which grows whatever
    # capacity an input needs and re-runs; (2) passes over all sets through the timed (asynchronous, CUDA-graph) path until a
    # whole pass changes nothing -- no capacity growth pending or done, no graph instantiated, no re-sort -- so that nothing
    # but the evaluation itself can fall into the timed region
    sampler = ClockSampler(local)
    for k in range(JITTER_SETS):
        one_eval(k, want_energy=True)
    settle_passes = 0
    n_warm = 0
    while True:
        st0 = stats()
        f0 = faults[0]
        for k in range(max(JITTER_SETS, 5)):
            flush.zero_()
            one_eval(k)
            n_warm += 1
        sync()
        st1 = stats()
        settle_passes += 1
        changed = float(np.any(st0[:7] != st1[:7]) or st1[7] != 0 or faults[0] != f0)
        if world > 1:
            tch = torch.tensor([changed], dtype=torch.float64, device=dev)
            dist.all_reduce(tch, op=dist.ReduceOp.MAX)
            changed = float(tch.item())
        if (not changed and n_warm >= max(args.warmup, 3)) or settle_passes >= 8:
            break
    stats_before = stats()
    tolerant[0] = False

    # ---- timed region: K steps of the PRODUCT path (one CUDA-graph launch per evaluation at N = 1; no profiling hooks),
    # one CUDA-event bracket per step on the launching stream, L2 flushed between steps outside the brackets
    K = args.steps
    launches0 = L.agbnp_b200_launch_count(handle)
    d_force.zero_(); d_energy.zero_()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    sync()
    coll0 = ev.collectives if sharded else 0
    t_region0 = time.time()
    t_wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()
        ev0[k].record(stream)
        one_eval(k)
        ev1[k].record(stream)
    sync()
    t_wall = time.perf_counter() - t_wall0
    t_region1 = time.time()
    clocks = sampler.result(t_region0, t_region1)
    stats_after = stats()
    coll_per_step = ((ev.collectives - coll0) / K) if sharded else 0
    launches = L.agbnp_b200_launch_count(handle) - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in zip(ev0, ev1)])
    tt = torch.tensor(np.concatenate([[step_ms.sum()], step_ms]), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)       # every step, and the total, as the slowest rank saw it
    dev_ms = float(tt[0].item())
    step_ms = tt[1:].cpu().numpy()
    ms_per_step = dev_ms / K
    value = (1e3 / ms_per_step) * (world if (world > 1 and not sharded) else 1)
    e_mean = float(d_energy.item()) / K
    step_stats = {"median": float(np.median(step_ms)), "p10": float(np.percentile(step_ms, 10)), "p90": float(np.percentile(step_ms, 90)),
                  "min": float(step_ms.min()), "max": float(step_ms.max()), "mean": float(step_ms.mean()),
                  "evals_per_s_at_median": 1e3 / float(np.median(step_ms))}
    settle = {"sync_evals": JITTER_SETS, "async_warmup_evals": n_warm, "passes": settle_passes,
              "during_timed_region": {"capacity_growths": int(stats_after[0]-stats_before[0]), "re_sorts": int(stats_after[1]-stats_before[1]),
                                      "graph_instantiations": int(stats_after[2]-stats_before[2]), "async_faults": int(stats_after[3]-stats_before[3])},
              "capacities": {"nodes_per_root": int(stats_after[4]), "nodes_per_level": int(stats_after[5]), "level2_neighbors": int(stats_after[6])}}

    # ---- end to end through the plugin interface with HOST buffers (single-GPU handle; sharded: the evaluator): per step, host
    # clock around the call alone (H2D of the step's positions from pinned memory, evaluation, D2H of forces + energy, all inside;
    # the L2 flush between steps is issued and waited for outside the clock, as in the device-resident loop)
    e2e = None
    if not sharded:
        host_sets = [jittered(s["pos"], k) for k in range(JITTER_SETS)]
        for k in range(3):
            ctx.setPositions(host_sets[k]); ctx.calcForcesAndEnergy()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = 0.0
        for k in range(K):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.setPositions(host_sets[k % JITTER_SETS])
            ctx.calcForcesAndEnergy()                   # synchronous: returns with forces and energy on the host
            dt += time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": K * (world if world > 1 else 1) / dt, "unit": UNIT, "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 12 * n + 560,
               "how": "AGBNPplugin Context.setPositions + calcForcesAndEnergy (agbnp_b200_execute_host): pinned staging, H2D positions, D2H forces+energy, host clock around each call, L2 flushed between calls"}
    else:
        pinned = [p.cpu().pin_memory() for p in posq_sets]
        h_force = torch.zeros((n, 3), dtype=torch.float32).pin_memory()
        d_in = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        sync()
        dt = 0.0
        for k in range(K):
            flush.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            if rank == 0:
                d_in.copy_(pinned[k % JITTER_SETS], non_blocking=True)
            d_force.zero_()
            ev.evaluate(d_in, sp, d_force, 0, n, None, True)
            if rank == 0:
                h_force.copy_(d_force, non_blocking=True)
            torch.cuda.synchronize()
            dt += time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": K / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 12 * n + 8,
               "how": "rank 0: pinned H2D positions -> broadcast -> sharded evaluation -> D2H forces + energy, host clock around each step, L2 flushed between steps"}

    # ---- per-kernel breakdown (separate, untimed pass: the profiling hooks bracket every launch, which needs the plain-launch
    # path) and the work counters
    sums = (C.c_double * 16)(); cnts = (C.c_int * 16)(); names_p = C.c_char_p()
    L.agbnp_b200_profile(handle, 0xffffffff)
    for k in range(24):
        flush.zero_()
        one_eval(k)
    nk = L.agbnp_b200_profile_read(handle, sums, cnts, 16, C.byref(names_p))
    L.agbnp_b200_profile(handle, 0)
    kn = names_p.value.decode().split("\n")
    kernels_us = {kn[i]: round(sums[i] / max(1, cnts[i]) * 1e3, 2) for i in range(nk) if cnts[i] > 0}
    wc = np.zeros(8)
    L.agbnp_b200_get(handle, _lib.GET["WORK_COUNTERS"], wc.ctypes.data_as(C.c_void_p), wc.nbytes)
    if world > 1:
        tw = torch.from_numpy(wc).to(dev)
        dist.all_reduce(tw)
        wc = tw.cpu().numpy()
    p_gb, p_q, c2, c3, m_nodes = wc[0], wc[1], wc[2], wc[3], wc[4]
    # algorithmic work of every kernel (SURVEY 8d per-unit figures; the kernels' shares add up to the path total)
    work = {"k_tree": (FLOP_CAND * (c2 + c3) + 150.0 * m_nodes, MUFU_CAND * (c2 + c3) + MUFU_NODE * m_nodes),     # build 25 + rescan 45 + two sweeps 80
            "k_born": (28.0 * p_q, 1.0 * p_q), "k_gb": (FLOP_GB * p_gb, MUFU_GB * p_gb), "k_deriv": (58.0 * p_q, 2.0 * p_q),
            "k_tree_gamma": (40.0 * m_nodes, 0.0)}
    flop = sum(w[0] for w in work.values())
    mufu = sum(w[1] for w in work.values())
    t_roof_ms = max(flop / (peak_fp32 * world), mufu / (peak_mufu * world)) * 1e3
    ktot = sum(kernels_us.values())
    table = {}
    for name, us in kernels_us.items():
        fl, mu = work.get(name, (0.0, 0.0))
        troof = max(fl / world / peak_fp32, mu / world / peak_mufu) * 1e6
        table[name] = {"us": us, "share": round(us / ktot, 4), "algorithmic_gflop": round(fl / world / 1e9, 4), "algorithmic_mufu_m": round(mu / world / 1e6, 3),
                       "t_roof_us": round(troof, 2), "frac": round(troof / us, 4) if us > 0 else None}
    dom = max(kernels_us, key=kernels_us.get)          # the kernel with the largest share of the evaluation
    d_fl, d_mu = work.get(dom, (0.0, 0.0))
    d_ms = kernels_us[dom] * 1e-3
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        traffic = {}
    roofline = {"bound": "fp32", "kernel": dom, "achieved": d_fl / world / (d_ms * 1e-3) / 1e12, "peak": peak_fp32 / 1e12, "unit": "TFLOP/s",
                "frac": None, "traffic": (traffic.get(dom) or {}).get("dram_bytes_per_launch"),
                "peak_source": "measured here by agbnp_b200_measure_peaks (FFMA / FFMA2 issue rate; MEASURED_PEAKS.json has no FP32 figure)",
                "launch_ms": d_ms, "algorithmic_flop_per_launch": d_fl / world, "share_of_step": kernels_us[dom] / ktot,
                "mufu_achieved_gops": d_mu / world / (d_ms * 1e-3) / 1e9, "mufu_peak_gops": peak_mufu / 1e9,
                "note": "dominant kernel by time share in the per-kernel pass; roofline_kernels has every kernel, path_roofline the whole evaluation"}
    roofline["frac"] = max(roofline["achieved"] / roofline["peak"], roofline["mufu_achieved_gops"] / roofline["mufu_peak_gops"])
    path_roofline = {"flop": flop, "mufu": mufu, "t_roof_ms": t_roof_ms, "t_eval_ms": ms_per_step, "frac": t_roof_ms / ms_per_step,
                     "frac_at_median": t_roof_ms / step_stats["median"],
                     "peak_fp32_tflops": peak_fp32 / 1e12, "peak_ffma2_tflops": 2 * pk[1] / 1e12, "peak_mufu_gops": peak_mufu / 1e9,
                     "peaks_measured": {"ffma_tflops": 2 * pk[0] / 1e12, "ffma2_tflops": 2 * pk[1] / 1e12, "mufu_ex2_gops": pk[2] / 1e9, "mufu_rsq_gops": pk[3] / 1e9,
                                        "theoretical_fp32_tflops_at_sm_clock": (148 * 128 * 2 * clocks["sm_mhz"] * 1e6 / 1e12) if clocks.get("sm_mhz") else None,
                                        "theoretical_mufu_gops_at_sm_clock": (148 * 16 * clocks["sm_mhz"] * 1e6 / 1e9) if clocks.get("sm_mhz") else None},
                     "counters": {"P_gb": p_gb, "P_q": p_q, "C2": c2, "C3plus": c3, "M": m_nodes}}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    tree_ms = kernels_us.get("k_tree", 0.0) * 1e-3
    tree_traffic = (traffic.get("k_tree") or {}).get("dram_bytes_per_launch")
    tree_bytes = TREE_STORE_BYTES_PER_NODE * m_nodes / world     # what must reach HBM: the persisted per-node store
    roofline_tree = {"bound": "hbm", "kernel": "k_tree", "achieved": tree_bytes / max(tree_ms, 1e-9) / 1e6, "peak": hbm_peak,
                     "unit": "GB/s", "frac": tree_bytes / max(tree_ms, 1e-9) / 1e6 / hbm_peak, "traffic": tree_traffic, "peak_source": hbm_src,
                     "algorithmic_bytes_per_launch": tree_bytes,
                     "note": "HBM view of the tree build (north_star: achieved HBM GB/s for the tree passes): algorithmic bytes = the persisted node store; "
                             "traffic = dram bytes of one launch from the committed ncu capture"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if (sharded or world == 1) else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic" if s.get("name", "").startswith("hivrt-standin") else "example/hivrt_agbnp1.dms",
            "config": config_dict(s, args, "1 GPU" if world == 1 else ("one evaluation sharded over %d GPUs, %s" % (world, "peer-memory exchanges and position broadcast over NVLink" if getattr(sk, "peer", False) else "NCCL all-reduces") if sharded else "%d independent replicas" % world)),
            "ns_per_day": value * NS_PER_DAY_PER_EVAL_PER_S, "roofline": roofline, "path_roofline": path_roofline, "roofline_tree": roofline_tree,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "kernels_us": kernels_us, "roofline_kernels": table,
            "step_ms": step_stats, "settle": settle, "mean_energy_kj_mol": e_mean, "host_wall_ms_per_step": t_wall / K * 1e3,
            "timed_path": ("one CUDA-graph launch per evaluation (%d kernels), asynchronous agbnp_b200_execute_device" % (launches // max(K, 1))) if not sharded
                          else "agbnp_b200_shard_evaluate: %d kernel launches per evaluation incl. exchanges" % (launches // max(K, 1))}
    if s.get("name", "").startswith("hivrt-standin"):
        line["config"]["stand_in"] = "example/hivrt_agbnp1.dms is absent from the reference checkout (.MISSING_LARGE_BLOBS); 2clr x 3 stand-in, N=17949 (SURVEY 8d)"
    if sharded:
        line["collectives_per_step"] = coll_per_step
        # the other way to use N GPUs at this size (BASELINE config 5): N independent replicas, one handle per GPU, no
        # communication -- reported next to the sharded number, never instead of it
        rctx = plug.Context(force, device=local)
        rh = rctx.kernel.handle

        def rep_eval(k, strict=True):
            rc = L.agbnp_b200_execute_device(rh, posq_sets[k % JITTER_SETS].data_ptr(), sp, d_force.data_ptr(), 0, n, None, None)
            if rc != 0 and (strict or rc != _lib.ERR_CAPACITY):
                raise RuntimeError(L.agbnp_b200_last_error(rh).decode())
        e = C.c_double(0.0)
        for k in range(JITTER_SETS):                # settle the capacities on every input (synchronous path)
            L.agbnp_b200_execute_device(rh, posq_sets[k].data_ptr(), sp, d_force.data_ptr(), 0, n, None, C.byref(e))
        for k in range(2*JITTER_SETS):
            rep_eval(k, strict=False)
        L.agbnp_b200_synchronize(rh, sp); dist.barrier(); torch.cuda.synchronize()
        r0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        r1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for k in range(K):
            flush.zero_()
            r0[k].record(stream); rep_eval(k); r1[k].record(stream)
        L.agbnp_b200_synchronize(rh, sp); dist.barrier(); torch.cuda.synchronize()
        tr = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(r0, r1))], dtype=torch.float64, device=dev)
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        line["replica_mode"] = {"value": world*K*1e3/float(tr.item()), "unit": UNIT, "ms_per_step": float(tr.item())/K, "scaling": "weak",
                                "what": "%d independent evaluations in flight, one per GPU (max over ranks of the per-GPU time)" % world}
        rctx.kernel.close()
    golden_parity = None
    if _WORKLOAD["method"] == 0 and not _WORKLOAD.get("tree_reuse"):
        # parity of the path that was just timed (sharded included) against the COMMITTED outputs of the compiled reference
        # (tests/golden/ref_outputs_large.npz, tools/make_golden.py) on the unjittered positions: cheap, so every line has it
        gname = "hivrt_standin" if s.get("name", "").startswith("hivrt-standin") else s.get("name", "")
        try:
            gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs_large.npz"))
            gold = {"e": float(gold[gname + "_v1_energy"]), "f": gold[gname + "_v1_forces"], "m": int(gold[gname + "_v1_tree_size"])} if gname + "_v1_energy" in gold.files else None
        except Exception:
            gold = None
        if gold is not None:
            p = torch.zeros((n, 4), dtype=torch.float32)
            p[:, :3] = torch.from_numpy(s["pos"].astype(np.float32))
            p = p.to(dev)
            d_force.zero_()
            if sharded:
                e_g = ev.evaluate(p, sp, d_force, 0, n, None, True)
            else:
                e_c = C.c_double(0.0)
                rc = L.agbnp_b200_execute_device(handle, p.data_ptr(), sp, d_force.data_ptr(), 0, n, None, C.byref(e_c))
                assert rc == 0, L.agbnp_b200_last_error(handle).decode()
                e_g = e_c.value
            torch.cuda.synchronize()
            f_g = d_force.cpu().numpy().astype(np.float64)
            tsz = np.zeros(1, dtype=np.int64)
            L.agbnp_b200_get(handle, _lib.GET["TREE_SIZE"], tsz.ctypes.data_as(C.c_void_p), tsz.nbytes)
            nodes = float(tsz[0])
            if world > 1 and sharded:
                tn = torch.tensor([nodes], dtype=torch.float64, device=dev)
                dist.all_reduce(tn)
                nodes = float(tn.item())
            golden_parity = {"energy_rel_err": abs(e_g - gold["e"]) / abs(gold["e"]),
                             "force_rel_rms": float(np.sqrt(((f_g - gold["f"]) ** 2).sum() / (gold["f"] ** 2).sum())),
                             "tree_nodes": int(nodes), "tree_nodes_reference": gold["m"] - 1 - n,
                             "against": "committed outputs of the compiled reference (tests/golden/ref_outputs_large.npz)",
                             "tolerance": {"energy_rel": 1e-5, "force_rel_rms": 1e-4, "tree_nodes": "equal"}}
            golden_parity["ok"] = bool(golden_parity["energy_rel_err"] <= 1e-5 and golden_parity["force_rel_rms"] <= 1e-4 and
                                       golden_parity["tree_nodes"] == golden_parity["tree_nodes_reference"])
            line["parity"] = golden_parity
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # unjittered evaluation through the synchronous device path for the full-size parity check
        p = torch.zeros((n, 4), dtype=torch.float32)
        p[:, :3] = torch.from_numpy(s["pos"].astype(np.float32))
        p = p.to(dev)
        d_force.zero_()
        e = C.c_double(0.0)
        rc = L.agbnp_b200_execute_device(handle, p.data_ptr(), sp, d_force.data_ptr(), 0, n, None, C.byref(e))
        assert rc == 0
        torch.cuda.synchronize()
        base, parity = cpu_baseline(s, e.value, d_force.cpu().numpy().astype(np.float64))
        line["cpu_baseline"] = base
        line["parity_live"] = parity
        if "parity" not in line:
            line["parity"] = parity
    if rank == 0:
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def md_arm(args):
    """bench.py --md STEPS [--workload W --cutoff C]: the reference's MD-benchmark protocol (example/hivrt_benchmark.py:17-33:
    LangevinIntegrator(300 K, 1/ps, 1 fs), simulation.step(N)) driven on the GPU through md.LangevinMD -- per step one
    asynchronous AGBNP evaluation and one fused integrator kernel.  Unlike the evaluation benchmark (independent jittered
    inputs), consecutive steps here are a trajectory: the Verlet lists (pair masks, level-2 candidates) are rebuilt when
    atoms have moved, as in production.  One JSON line: ns/day (BASELINE configs 2-3) next to steps/s."""
    import torch
    from openmm_agbnp_plugin_b200 import systems, md
    s = workload()
    n = len(s["pos"])
    force = systems.make_force(s, 1, _WORKLOAD["method"], _WORKLOAD["cutoff"])
    masses = np.where(s["ishydrogen"] > 0, 1.008, 12.0)
    sim = md.LangevinMD(force, s["pos"], masses, temperature=300.0, friction_per_ps=args.md_friction, dt_ps=0.001, restraint_k=args.md_tether, seed=20261018)
    warm = max(200, args.warmup)
    sim.step(warm)                      # thermalise: lists get rebuilt at the rate of a running simulation
    sim.synchronize()
    st0 = sim.stats()
    dropped0 = sim.dropped
    sampler = ClockSampler(0)
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_r0 = time.time()
    e0.record(stream)
    sim.step(args.md)
    e1.record(stream)
    sim.synchronize()
    torch.cuda.synchronize()
    t_r1 = time.time()
    clocks = sampler.result(t_r0, t_r1)
    st1 = sim.stats()
    ms = e0.elapsed_time(e1) / args.md
    temp = sim.temperature()
    ls = sim.list_stats()
    # per-kernel times at the end state of the trajectory (separate, untimed: the profiling hooks need the plain-launch path)
    from openmm_agbnp_plugin_b200 import _lib
    L = _lib.lib()
    hnd = sim.kernel.handle
    sums = (C.c_double * 16)(); cnts = (C.c_int * 16)(); names_p = C.c_char_p()
    L.agbnp_b200_profile(hnd, 0xffffffff)
    sim.step(40)
    sim.synchronize()
    nk = L.agbnp_b200_profile_read(hnd, sums, cnts, 16, C.byref(names_p))
    L.agbnp_b200_profile(hnd, 0)
    kn = names_p.value.decode().split("\n")
    kernels_us = {kn[i]: round(sums[i] / max(1, cnts[i]) * 1e3, 2) for i in range(nk) if cnts[i] > 0}
    line = {"metric": "AGBNP1 MD throughput", "value": (1e3 / ms) * NS_PER_DAY_PER_EVAL_PER_S, "unit": "ns/day",
            "steps_per_s": 1e3 / ms, "ms_per_step": ms, "n_gpus": 1, "steps": args.md, "warmup": warm, "higher_is_better": True, "dtype": "f32",
            "data": s.get("name", _WORKLOAD["name"]), "config": config_dict(s, args, "1 GPU, Langevin 300 K, %g/ps, dt 1 fs, AGBNP only + %g kJ/mol/nm^2 tether" % (args.md_friction, args.md_tether)),
            "temperature_K": temp, "clocks": clocks, "kernels_us": kernels_us,
            "capacities": {"nodes_per_root": int(st1[4]), "nodes_per_level": int(st1[5]), "level2_neighbors": int(st1[6])},
            "during_timed_region": {"capacity_growths": int(st1[0]-st0[0]), "re_sorts": int(st1[1]-st0[1]), "async_faults": int(st1[3]-st0[3]),
                                    "evaluations_not_delivered": int(sim.dropped-dropped0)},
            "verlet_lists": {"skin_nm": float(ls[3]), "evaluations_since_voided": int(ls[2]), "pair_mask_rebuilds": int(ls[0]), "level2_list_rebuilds": int(ls[1])},
            "note": "one AGBNP evaluation + one integrator kernel per step, nothing on the host; device time by CUDA events over all steps"}
    line["config"]["inputs"] = "a Langevin trajectory (consecutive steps), not independent jittered inputs"
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    sim.close()


def main():
    # stdout carries exactly ONE JSON line: libraries that print banners there (NCCL's version line) go to stderr instead
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="shard", choices=["shard", "replica"], help="N>1: shard one evaluation (default) or run replicas")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="hivrt", help="hivrt (the metric's workload, default) | 2clr | 1dwc | rnaseh | 1li2 | trpcage")
    ap.add_argument("--cutoff", type=float, default=0.0, help="> 0: CutoffNonPeriodic with this cutoff (nm); the Reference platform, and "
                    "therefore the CPU baseline / parity check, has no cutoff: they are skipped")
    ap.add_argument("--md", type=int, default=0, help="> 0: MD-driven throughput instead (this many Langevin steps at 1 fs through md.LangevinMD, "
                    "the reference's example/*_benchmark.py protocol); prints ns/day for --workload / --cutoff")
    ap.add_argument("--md-tether", type=float, default=100000.0, help="harmonic tether of every atom to its start position (kJ/mol/nm^2) in --md: stands in "
                    "for the bonded and repulsive terms AGBNP does not have; 1e5 gives 0.005 nm rms per coordinate at 300 K, i.e. contact "
                    "distances that fluctuate about as in a protein")
    ap.add_argument("--md-friction", type=float, default=10.0, help="Langevin friction (1/ps) of --md.  The reference scripts use 1/ps with the "
                    "full force field; with AGBNP alone and CutoffNonPeriodic (pair terms truncated without switching, by the reference's "
                    "definition) 1/ps does not hold 300 K -- the truncation heats -- so the default is 10/ps; the line reports the temperature")
    ap.add_argument("--tree-reuse", type=int, default=0, help="> 1: OPT-IN tree reuse (not the reference's semantics): the overlap tree is "
                    "rebuilt every K-th evaluation and re-evaluated on its stored topology in between; reported in config")
    args = ap.parse_args()
    if args.tree_reuse > 1:
        os.environ["AGBNP_B200_TREE_REUSE"] = str(args.tree_reuse)
        _WORKLOAD["tree_reuse"] = args.tree_reuse
    _WORKLOAD["name"] = args.workload
    if args.cutoff > 0:
        _WORKLOAD["method"], _WORKLOAD["cutoff"] = 1, args.cutoff
        args.no_cpu_baseline = True
    if args.impl == "reference":
        reference_arm(args)
    elif args.md > 0:
        md_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
