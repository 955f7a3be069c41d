"""ctypes binding of lib/libagbnp_b200.so (C-ABI: include/agbnp_b200.h).  Fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AGBNP_B200_LIB: another build of the same library (tuning experiments: tools/build_variant.sh); never a fallback
LIB_PATH = os.environ.get("AGBNP_B200_LIB") or os.path.join(_HERE, "lib", "libagbnp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "agbnp_b200.h")

OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_PARAM_CHANGE = 0, -1, -2, -3, -4
NOCUTOFF, CUTOFF_NONPERIODIC, CUTOFF_PERIODIC = 0, 1, 2

GET = dict(SELF_VOLUME_VDW=0, SELF_VOLUME_LARGE=1, SURFACE_AREA=2, BORN_RADIUS=3, VOLUME_SCALING=4, SCALARS=5,
           TREE_SIZE=6, TREE_TOPOLOGY=7, DERIV_Y=8, DERIV_WU=9, NEIGHBOR_PAIRS=10, NEIGHBOR_COUNT=11, WORK_COUNTERS=12, STATS=13, LIST_STATS=14, PEER_STATE=15)
BUF = dict(SELFVOL=0, YQ=1, FORCE=2, ENERGY=3, WU=4, BSUM=5)


class Config(C.Structure):
    _fields_ = [("version", C.c_int), ("nonbonded_method", C.c_int), ("cutoff", C.c_double), ("device", C.c_int),
                ("shard_rank", C.c_int), ("shard_count", C.c_int), ("reorder_interval", C.c_int),
                ("tree_reuse_interval", C.c_int)]


class DeviceLayout(C.Structure):
    _fields_ = [("atom_index", C.POINTER(C.c_int)), ("posq_is_double", C.c_int), ("energy_is_float", C.c_int)]


_lib = None


def lib():
    """Load the CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libagbnp_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(%s missing; there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, dp, ucp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_ubyte)
        L.agbnp_b200_version.restype = C.c_char_p
        L.agbnp_b200_default_config.argtypes = [C.POINTER(Config)]
        L.agbnp_b200_default_config.restype = None
        L.agbnp_b200_create.argtypes = [C.POINTER(Config), C.c_int, dp, dp, dp, dp, ucp, C.POINTER(vp)]
        L.agbnp_b200_destroy.argtypes = [vp]
        L.agbnp_b200_destroy.restype = None
        L.agbnp_b200_last_error.argtypes = [vp]
        L.agbnp_b200_last_error.restype = C.c_char_p
        L.agbnp_b200_set_params.argtypes = [vp, C.c_int, dp, dp, dp, dp, ucp]
        L.agbnp_b200_execute_host.argtypes = [vp, dp, C.c_int, C.c_int, dp, dp]
        L.agbnp_b200_execute_device.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, dp]
        L.agbnp_b200_synchronize.argtypes = [vp, vp]
        L.agbnp_b200_set_device_layout.argtypes = [vp, C.POINTER(DeviceLayout)]
        L.agbnp_b200_profile.argtypes = [vp, C.c_uint]
        L.agbnp_b200_profile_read.argtypes = [vp, dp, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_char_p)]
        L.agbnp_b200_launch_count.argtypes = [vp]
        L.agbnp_b200_launch_count.restype = C.c_longlong
        L.agbnp_b200_measure_peaks.argtypes = [C.c_int, dp, C.c_int]
        L.agbnp_b200_get.argtypes = [vp, C.c_int, vp, C.c_size_t]
        L.agbnp_b200_shard_phase.argtypes = [vp, C.c_int, vp, vp]
        L.agbnp_b200_shard_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.agbnp_b200_shard_finish.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, dp]
        L.agbnp_b200_peer_export.argtypes = [vp, vp]
        L.agbnp_b200_peer_import.argtypes = [vp, vp, C.c_int]
        L.agbnp_b200_peer_import_local.argtypes = [vp, C.POINTER(vp), C.c_int]
        L.agbnp_b200_peer_exchange.argtypes = [vp, C.c_int, vp]
        L.agbnp_b200_peer_broadcast.argtypes = [vp, vp, C.c_int, vp]
        L.agbnp_b200_shard_evaluate.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp]
        _lib = L
    return _lib


def declared_symbols():
    """Function names declared in include/agbnp_b200.h (used by the ABI test)."""
    import re
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agbnp_b200_[a-z_0-9]+)\s*\(", text)))
