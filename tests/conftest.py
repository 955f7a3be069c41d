import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # the oracle's C restatement is test infrastructure: make sure it is built before anything imports it
    pass


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    import subprocess
    so = os.path.join(ROOT, "oracle", "libagbnp_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    yield


def load_system(name):
    if name == "gaussvol":
        s = np.load(os.path.join(GOLDEN, "gaussvol.npz"))
    else:
        s = np.load(os.path.join(GOLDEN, "systems", name + ".npz"))
    return {k: s[k] for k in s.files}


def sys_args(s):
    return (s["radius"], s["gamma"], s["alpha"], s["charge"], s["ishydrogen"])


def relrms(a, b):
    return float(np.sqrt(((a - b) ** 2).sum() / max(float((b ** 2).sum()), 1e-300)))


def gpu_topology(rows):
    """agbnp_b200 TREE_TOPOLOGY dump -> {parent path: [child atoms in sibling order]}"""
    m = len(rows)
    paths = [None] * m
    kids = {}
    for w in range(m):
        root, par, atom, rank = (int(x) for x in rows[w])
        ppath = (root,) if par < 0 else paths[par]
        paths[w] = ppath + (atom,)
        kids.setdefault(ppath, []).append((rank, atom))
    return {p: [a for _, a in sorted(v)] for p, v in kids.items()}


@pytest.fixture(scope="session")
def golden():
    import json
    return json.load(open(os.path.join(GOLDEN, "golden.json")))


@pytest.fixture(scope="session")
def ref_outputs():
    return np.load(os.path.join(GOLDEN, "ref_outputs.npz"))


@pytest.fixture(scope="session")
def ref_large():
    """Outputs of the compiled, unmodified reference on 1dwc, 2clr and the full-size HIV-RT stand-in (tools/make_golden.py)."""
    return np.load(os.path.join(GOLDEN, "ref_outputs_large.npz"))


def pair_keys(pairs, n):
    """(i<j) pair list -> sorted int64 keys, for set comparison of million-pair lists without Python sets"""
    p = np.asarray(pairs, dtype=np.int64)
    return np.sort(p[:, 0] * n + p[:, 1])
