"""ctypes binding of oracle/_ref/libagbnp_ref.so (the compiled, unmodified reference).  Test infrastructure only."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libagbnp_ref.so")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libagbnp_ref.so not built (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(LIB_PATH)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_check_version.argtypes = [C.c_int]
        L.ref_create.restype = vp
        L.ref_create.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, ip]
        L.ref_destroy.argtypes = [vp]
        L.ref_set_params.argtypes = [vp, dp, dp, dp, dp, ip]
        L.ref_execute.argtypes = [vp, dp, dp, dp]
        L.ref_get.argtypes = [vp, C.c_int, dp]
        L.ref_i4_dims.argtypes = [vp, ip, ip, ip]
        L.ref_i4_table.argtypes = [vp, C.c_int, C.c_int, dp, dp, dp]
        L.ref_i4_eval.argtypes = [vp, C.c_double, C.c_int, C.c_int, dp, dp]
        L.ref_kernel_tree_size.argtypes = [vp]
        L.gv_create.restype = vp
        L.gv_create.argtypes = [C.c_int, ip]
        L.gv_destroy.argtypes = [vp]
        L.gv_update.argtypes = [vp, C.c_int, dp, dp, dp, dp]
        L.gv_volume.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp]
        L.gv_tree_size.argtypes = [vp]
        L.gv_tree_dump.argtypes = [vp, ip, ip, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp]
        L.ref_constants.argtypes = [dp]
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


class ReferenceError_(RuntimeError):
    pass


def _err():
    return ReferenceError_(lib().ref_last_error().decode())


def constants():
    out = np.zeros(8)
    lib().ref_constants(out.ctypes.data_as(C.POINTER(C.c_double)))
    return dict(KFC=out[0], VOLMINA=out[1], VOLMINB=out[2], MIN_GVOL=out[3], RADIUS_INCREMENT=out[4],
                HB_RADIUS=out[5], I4LOOKUP_MAXA=out[6], MAX_ORDER=int(out[7]))


def check_version(v):
    if lib().ref_check_version(int(v)) != 0:
        raise _err()


class ReferenceKernel:
    """ReferenceCalcAGBNPForceKernel behind AGBNPForce (ReferenceAGBNPKernels.cpp:58-149)."""

    def __init__(self, version, radius, gamma, alpha, charge, ishydrogen):
        self.n = len(radius)
        self._keep = [_d(radius), _d(gamma), _d(alpha), _d(charge), _i(ishydrogen)]
        self.h = lib().ref_create(int(version), self.n, *[k[1] for k in self._keep])
        if not self.h:
            raise _err()

    def close(self):
        if self.h:
            lib().ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, radius, gamma, alpha, charge, ishydrogen):
        k = [_d(radius), _d(gamma), _d(alpha), _d(charge), _i(ishydrogen)]
        if lib().ref_set_params(self.h, *[x[1] for x in k]) != 0:
            raise _err()

    def execute(self, pos):
        p, pp = _d(pos)
        e = C.c_double(0)
        f = np.zeros((self.n, 3))
        if lib().ref_execute(self.h, pp, C.byref(e), f.ctypes.data_as(C.POINTER(C.c_double))) != 0:
            raise _err()
        return e.value, f

    def get(self, what):
        sel = dict(self_volume=0, volume_scaling_factor=1, born_radius=2, inverse_born_radius_fp=3,
                   radius_type_screened=4, radius_type_screener=5)[what]
        out = np.zeros(self.n)
        if lib().ref_get(self.h, sel, out.ctypes.data_as(C.POINTER(C.c_double))) != 0:
            raise _err()
        return out

    def i4_tables(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().ref_i4_dims(self.h, C.byref(a), C.byref(b), C.byref(c))
        ni, nj, nn = a.value, b.value, c.value
        x = np.zeros((ni, nj, nn)); y = np.zeros_like(x); y2 = np.zeros_like(x)
        dp = C.POINTER(C.c_double)
        for i in range(ni):
            for j in range(nj):
                lib().ref_i4_table(self.h, i, j, x[i, j].ctypes.data_as(dp), y[i, j].ctypes.data_as(dp),
                                   y2[i, j].ctypes.data_as(dp))
        return x, y, y2

    def i4_eval(self, d, ti, tj):
        q, dq = C.c_double(), C.c_double()
        if lib().ref_i4_eval(self.h, float(d), int(ti), int(tj), C.byref(q), C.byref(dq)) != 0:
            raise _err()
        return q.value, dq.value

    def tree_size(self):
        return lib().ref_kernel_tree_size(self.h)


class GaussVolTree:
    """Direct GaussVol access (gaussvol.h:205-312): build / rescan / up-sweep and a flat tree dump."""

    def __init__(self, ishydrogen):
        self.n = len(ishydrogen)
        self._ish = _i(ishydrogen)
        self.h = lib().gv_create(self.n, self._ish[1])

    def close(self):
        if self.h:
            lib().gv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _update(self, mode, pos, radii, volumes, gammas):
        a = [_d(pos), _d(radii), _d(volumes), _d(gammas)]
        if lib().gv_update(self.h, mode, *[x[1] for x in a]) != 0:
            raise _err()

    def compute_tree(self, pos, radii, volumes, gammas):
        self._update(0, pos, radii, volumes, gammas)

    def rescan_volumes(self, pos, radii, volumes, gammas):
        self._update(1, pos, radii, volumes, gammas)

    def rescan_gammas(self, pos, radii, volumes, gammas):
        self._update(2, pos, radii, volumes, gammas)

    def compute_volume(self, pos):
        p, pp = _d(pos)
        n = self.n
        vol, en = C.c_double(), C.c_double()
        force = np.zeros((n, 3)); gradv = np.zeros(n); fv = np.zeros(n); sv = np.zeros(n)
        dp = C.POINTER(C.c_double)
        if lib().gv_volume(self.h, pp, C.byref(vol), C.byref(en), force.ctypes.data_as(dp), gradv.ctypes.data_as(dp),
                           fv.ctypes.data_as(dp), sv.ctypes.data_as(dp)) != 0:
            raise _err()
        return dict(volume=vol.value, energy=en.value, force=force, gradV=gradv, free_volume=fv, self_volume=sv)

    def dump(self):
        m = lib().gv_tree_size(self.h)
        ints = {k: np.zeros(m, dtype=np.int32) for k in ("level", "atom", "parent", "child_start", "child_count")}
        dbl = {k: np.zeros(m) for k in ("volume", "gvol", "gamma1i", "a", "dvv1", "sfp")}
        c = np.zeros((m, 3)); dv1 = np.zeros((m, 3))
        ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
        lib().gv_tree_dump(self.h, *[ints[k].ctypes.data_as(ip) for k in ("level", "atom", "parent", "child_start", "child_count")],
                           dbl["volume"].ctypes.data_as(dp), dbl["gvol"].ctypes.data_as(dp), dbl["gamma1i"].ctypes.data_as(dp),
                           dbl["a"].ctypes.data_as(dp), c.ctypes.data_as(dp), dv1.ctypes.data_as(dp),
                           dbl["dvv1"].ctypes.data_as(dp), dbl["sfp"].ctypes.data_as(dp))
        out = dict(ints); out.update(dbl); out["c"] = c; out["dv1"] = dv1
        return out
