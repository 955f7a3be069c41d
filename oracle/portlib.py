"""ctypes binding of oracle/libagbnp_oracle.so (plain-C restatement, oracle/agbnp_oracle.c).  Test infrastructure only."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libagbnp_oracle.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/libagbnp_oracle.so not built (run `make -C oracle port`)")
        L = C.CDLL(LIB_PATH)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.agbnp_oracle_last_error.restype = C.c_char_p
        L.agbnp_oracle_create.restype = vp
        L.agbnp_oracle_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, dp, dp, dp, dp, ip]
        L.agbnp_oracle_destroy.argtypes = [vp]
        L.agbnp_oracle_set_params.argtypes = [vp, dp, dp, dp, dp, ip]
        L.agbnp_oracle_execute.argtypes = [vp, dp, dp, dp]
        L.agbnp_oracle_get.argtypes = [vp, C.c_int, dp]
        L.agbnp_oracle_scalar.restype = C.c_double
        L.agbnp_oracle_scalar.argtypes = [vp, C.c_int]
        L.agbnp_oracle_counter.restype = C.c_double
        L.agbnp_oracle_counter.argtypes = [vp, C.c_int]
        L.agbnp_oracle_tree_size.argtypes = [vp]
        L.agbnp_oracle_tree_dump.argtypes = [vp, ip, ip, ip, ip, ip, dp, dp]
        L.agbnp_oracle_i4_dims.argtypes = [vp, ip, ip, ip]
        L.agbnp_oracle_i4_table.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp]
        L.agbnp_oracle_neighbor_pairs.restype = C.c_long
        L.agbnp_oracle_neighbor_pairs.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_float, ip, C.c_long]
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


class OracleError(RuntimeError):
    pass


def _err():
    return OracleError(lib().agbnp_oracle_last_error().decode())


NoCutoff, CutoffNonPeriodic, CutoffPeriodic = 0, 1, 2

_GET = dict(self_volume=0, self_volume_large=1, volume_scaling_factor=2, born_radius=3, inverse_born_radius_fp=4,
            Y=5, bru=6, brw=7, W=8, U=9, radius_type_screened=10, radius_type_screener=11, free_volume=12,
            free_volume_large=13)
_SCAL = dict(vol_energy1=0, vol_energy2=1, gb_self=2, gb_pair=3, evdw=4, volume1=5, volume2=6)
_CNT = dict(P_gb=0, P_q=1, C2=2, C3=3, M=4)


class OracleKernel:
    def __init__(self, version, radius, gamma, alpha, charge, ishydrogen, nonbonded_method=NoCutoff, cutoff=1.0):
        self.n = len(radius)
        k = [_d(radius), _d(gamma), _d(alpha), _d(charge), _i(ishydrogen)]
        self.h = lib().agbnp_oracle_create(int(version), int(nonbonded_method), float(cutoff), self.n, *[x[1] for x in k])
        if not self.h:
            raise _err()

    def close(self):
        if getattr(self, "h", None):
            lib().agbnp_oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, radius, gamma, alpha, charge, ishydrogen):
        k = [_d(radius), _d(gamma), _d(alpha), _d(charge), _i(ishydrogen)]
        if lib().agbnp_oracle_set_params(self.h, *[x[1] for x in k]) != 0:
            raise _err()

    def execute(self, pos):
        p, pp = _d(pos)
        e = C.c_double(0)
        f = np.zeros((self.n, 3))
        if lib().agbnp_oracle_execute(self.h, pp, C.byref(e), f.ctypes.data_as(C.POINTER(C.c_double))) != 0:
            raise _err()
        return e.value, f

    def get(self, what):
        out = np.zeros(self.n)
        if lib().agbnp_oracle_get(self.h, _GET[what], out.ctypes.data_as(C.POINTER(C.c_double))) != 0:
            raise _err()
        return out

    def scalar(self, what):
        return lib().agbnp_oracle_scalar(self.h, _SCAL[what])

    def counter(self, what):
        return lib().agbnp_oracle_counter(self.h, _CNT[what])

    def tree(self):
        m = lib().agbnp_oracle_tree_size(self.h)
        ints = {k: np.zeros(m, dtype=np.int32) for k in ("level", "atom", "parent", "child_start", "child_count")}
        vol = np.zeros(m); gvol = np.zeros(m)
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        lib().agbnp_oracle_tree_dump(self.h, *[ints[k].ctypes.data_as(ip) for k in ("level", "atom", "parent", "child_start", "child_count")],
                                     vol.ctypes.data_as(dp), gvol.ctypes.data_as(dp))
        out = dict(ints); out["volume"] = vol; out["gvol"] = gvol
        return out

    def i4_tables(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().agbnp_oracle_i4_dims(self.h, C.byref(a), C.byref(b), C.byref(c))
        ni, nj, nn = a.value, b.value, c.value
        x = np.zeros((ni, nj, nn)); y = np.zeros_like(x); y2 = np.zeros_like(x)
        dp = C.POINTER(C.c_double)
        for i in range(ni):
            for j in range(nj):
                lib().agbnp_oracle_i4_table(self.h, i, j, x[i, j].ctypes.data_as(dp), y[i, j].ctypes.data_as(dp),
                                            y2[i, j].ctypes.data_as(dp))
        return x, y, y2


def neighbor_pairs(pos_f32, cutoff):
    """All (i<j) with float r2 < cutoff2 (the membership rule shared with the GPU)."""
    p = np.ascontiguousarray(pos_f32, dtype=np.float32)
    n = p.shape[0]
    fp = p.ctypes.data_as(C.POINTER(C.c_float))
    cnt = lib().agbnp_oracle_neighbor_pairs(n, fp, C.c_float(cutoff), None, 0)
    out = np.zeros((cnt, 2), dtype=np.int32)
    lib().agbnp_oracle_neighbor_pairs(n, fp, C.c_float(cutoff), out.ctypes.data_as(C.POINTER(C.c_int)), cnt)
    return out


def tree_topology(tree):
    """Canonical topology: dict parent-path (tuple of atoms) -> list of child atoms in sibling order."""
    level, atom, parent = tree["level"], tree["atom"], tree["parent"]
    cs, cc = tree["child_start"], tree["child_count"]
    m = len(level)
    paths = [None]*m
    paths[0] = ()
    topo = {}
    for s in range(1, m):
        paths[s] = paths[parent[s]] + (int(atom[s]),)
    for s in range(1, m):
        if cc[s] > 0:
            topo[paths[s]] = [int(atom[c]) for c in range(cs[s], cs[s]+cc[s])]
    return topo
