"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/agbnp_b200.h declares, and its
argument validation reproduces the reference's error behaviour.  No compute calls (no GPU here); on a box without a
CUDA device a valid create must FAIL LOUDLY (there is no CPU fallback)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import load_system
from openmm_agbnp_plugin_b200 import _lib
import openmm_agbnp_plugin_b200 as plug
from openmm_agbnp_plugin_b200 import systems


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), "libagbnp_b200.so does not export %s" % n
    assert b"sm_100a" in L.agbnp_b200_version()


def test_library_is_sm100a_only():
    """The shipped .so carries sm_100a SASS only (no multi-arch fat binary, no PTX fallback for other chips)."""
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out
    archs = {w for line in out.splitlines() for w in line.replace(".", " ").split() if w.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_default_config_matches_reference_defaults():
    cfg = _lib.Config()
    _lib.lib().agbnp_b200_default_config(C.byref(cfg))
    # AGBNPForce.cpp:15
    assert (cfg.version, cfg.nonbonded_method, cfg.cutoff) == (1, _lib.NOCUTOFF, 1.0)
    assert (cfg.shard_rank, cfg.shard_count) == (0, 1)
    assert cfg.tree_reuse_interval == 0          # the opt-in tree reuse is off: the reference rebuilds every evaluation


def test_config_struct_matches_the_header():
    """agbnp_b200_config is passed by pointer: the ctypes mirror must list the header's fields in the header's order."""
    import re
    text = open(_lib.HEADER_PATH).read()
    body = text[text.index("typedef struct {"):text.index("} agbnp_b200_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(int|double)\s+(\w+)\s*;", body)
    ctype = {"int": C.c_int, "double": C.c_double}
    assert [(n, ctype[t]) for t, n in fields] == list(_lib.Config._fields_)


def _create(cfg_mod, s):
    L = _lib.lib()
    cfg = _lib.Config()
    L.agbnp_b200_default_config(C.byref(cfg))
    cfg_mod(cfg)
    dp = C.POINTER(C.c_double)
    arr = [np.ascontiguousarray(s[k], dtype=np.float64) for k in ("radius", "gamma", "alpha", "charge")]
    ish = np.ascontiguousarray(s["ishydrogen"], dtype=np.uint8)
    h = C.c_void_p()
    rc = L.agbnp_b200_create(C.byref(cfg), len(ish), *[a.ctypes.data_as(dp) for a in arr],
                             ish.ctypes.data_as(C.POINTER(C.c_ubyte)), C.byref(h))
    return rc, h, (L.agbnp_b200_last_error(None) or b"").decode()


def test_create_rejects_what_the_reference_rejects():
    s = load_system("trpcage")
    rc, h, msg = _create(lambda c: setattr(c, "version", 3), s)
    assert rc == _lib.ERR_ARG and "illegal version number" in msg          # AGBNPForce.cpp:52-59
    rc, h, msg = _create(lambda c: setattr(c, "version", 2), s)
    assert rc == _lib.ERR_ARG and "AGBNP2" in msg                            # out of this library's path
    rc, h, msg = _create(lambda c: setattr(c, "nonbonded_method", _lib.CUTOFF_PERIODIC), s)
    assert rc == _lib.ERR_ARG and "CutoffPeriodic" in msg                    # implemented nowhere in the reference
    rc, h, msg = _create(lambda c: setattr(c, "shard_count", 0), s)
    assert rc == _lib.ERR_ARG
    s2 = dict(s)
    g = s["gamma"].copy()
    heavy = np.where(s["ishydrogen"] == 0)[0]
    g[heavy[5]] += 1.0
    s2["gamma"] = g
    rc, h, msg = _create(lambda c: None, s2)
    assert rc == _lib.ERR_ARG
    assert msg == "initialize(): AGBNP does not support multiple gamma values."   # ReferenceAGBNPKernels.cpp:114


@pytest.mark.skipif(_cuda_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_gpu():
    s = load_system("trpcage")
    rc, h, msg = _create(lambda c: None, s)
    assert rc == _lib.ERR_CUDA and "no CPU fallback" in msg and not h.value
    with pytest.raises(plug.OpenMMException, match="no CPU fallback"):
        plug.Context(systems.make_force(s))


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or execute it."""
    import re
    pkg = os.path.dirname(_lib.__file__)
    bad = re.compile(r"(^|\s)(from|import)\s+oracle\b|libagbnp_oracle|libagbnp_ref|agbnp_oracle\.h|oracle/|dlopen|CDLL\([^)]*oracle")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                m = bad.search(text)
                assert m is None, "%s: %r" % (os.path.join(dirpath, f), m.group(0))
    import subprocess
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "oracle" not in ldd and "agbnp_ref" not in ldd
