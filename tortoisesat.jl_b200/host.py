"""ctypes binding of libtortoise_b200.so + reference-named host functions.

Reference interface mirrored here (file:line in /root/reference/src):
  igrf12(date, r, lat, lon)                       igrf.jl:67-274
  igrf_data(altitude, year)                       magnetic_toolbox.jl:108-127
(more entry points are added with each kernel)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TS_OK = 0
TS_ERR_CUDA, TS_ERR_ARG, TS_ERR_DATE, TS_ERR_DOMAIN, TS_ERR_NOMEM = -1, -2, -3, -4, -5

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class TortoiseError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("tortoise_b200 error %d: %s" % (code, msg))
        self.code = code


def lib_path():
    return os.path.join(_HERE, "libtortoise_b200.so")


def load_library():
    """Loads the CUDA library.  Fails loudly (no fallback) if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise RuntimeError(
            "libtortoise_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback." % p)
    L = C.CDLL(p)
    L.ts_version.restype = C.c_int
    L.ts_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.ts_create.restype = C.c_int
    L.ts_destroy.argtypes = [C.c_void_p]
    L.ts_destroy.restype = None
    L.ts_last_error.argtypes = [C.c_void_p]
    L.ts_last_error.restype = C.c_char_p
    L.ts_device_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_char_p, C.c_int]
    L.ts_launch_count.argtypes = [C.c_void_p]
    L.ts_launch_count.restype = C.c_int64
    L.ts_synchronize.argtypes = [C.c_void_p]
    L.ts_last_kernel_ms.argtypes = [C.c_void_p]
    L.ts_last_kernel_ms.restype = C.c_double
    L.ts_fp64_peak_probe.argtypes = [C.c_void_p, c_double_p]
    L.ts_igrf12_batch.argtypes = [C.c_void_p, C.c_double, C.c_int64] + [C.c_void_p] * 6 + [C.c_int]
    _LIB = L
    return L


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    """host numpy array or torch CUDA tensor -> raw address"""
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


class Engine:
    """One context per GPU (ts_ctx).  Not re-entrant."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.ts_create(C.byref(h), int(device))
        if rc != TS_OK:
            raise TortoiseError(rc, "ts_create(device=%d) failed: no usable CUDA device (there is no CPU fallback)" % device)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.ts_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != TS_OK:
            raise TortoiseError(rc, self.lib.ts_last_error(self.h).decode())

    # -- bookkeeping
    def device_info(self):
        sm = C.c_int()
        name = C.create_string_buffer(128)
        self._check(self.lib.ts_device_info(self.h, C.byref(sm), name, 128))
        return sm.value, name.value.decode()

    def launch_count(self):
        return int(self.lib.ts_launch_count(self.h))

    def last_kernel_ms(self):
        return float(self.lib.ts_last_kernel_ms(self.h))

    def synchronize(self):
        self._check(self.lib.ts_synchronize(self.h))

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._check(self.lib.ts_fp64_peak_probe(self.h, C.byref(v)))
        return v.value

    # -- K1 ---------------------------------------------------------------
    def igrf12_batch(self, date, r_m, lat, lon, out=None):
        """Batched igrf12 (igrf.jl:67-274).  numpy in -> numpy out (host path), or
        torch CUDA float64 tensors in -> `out` tensors filled (device path)."""
        if isinstance(r_m, np.ndarray) or np.isscalar(r_m) or isinstance(r_m, (list, tuple)):
            r_m, lat, lon = _f64(np.atleast_1d(r_m)), _f64(np.atleast_1d(lat)), _f64(np.atleast_1d(lon))
            n = r_m.shape[0]
            if not (lat.shape[0] == n and lon.shape[0] == n):
                raise ValueError("r, lat, lon must have the same length")
            Bn, Be, Bd = np.empty(n), np.empty(n), np.empty(n)
            rc = self.lib.ts_igrf12_batch(self.h, float(date), n, _ptr(r_m), _ptr(lat), _ptr(lon), _ptr(Bn), _ptr(Be),
                                          _ptr(Bd), 0)
            if rc == TS_ERR_DOMAIN:
                raise TortoiseError(rc, self.lib.ts_last_error(self.h).decode())
            self._check(rc)
            return Bn, Be, Bd
        n = r_m.numel()
        Bn, Be, Bd = out
        self._check(self.lib.ts_igrf12_batch(self.h, float(date), n, _ptr(r_m), _ptr(lat), _ptr(lon), _ptr(Bn), _ptr(Be),
                                             _ptr(Bd), 1))
        return Bn, Be, Bd


# ---------------------------------------------------------------------------
# Reference-named free functions (scalar calls route to batch-of-1; correctness
# path, not the performance path).  A module-level default engine is created on
# first use.
_DEFAULT = None


def default_engine():
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Engine(0)
    return _DEFAULT


def igrf12(date, r, lat, lon, show_warns=True):
    """igrf12(date, r, λ, Ω) -> [north, east, down] nT  (igrf.jl:67-274).
    Raises like the reference for date/lat/lon outside their ranges."""
    Bn, Be, Bd = default_engine().igrf12_batch(date, [r], [lat], [lon])
    return np.array([Bn[0], Be[0], Bd[0]])


def igrf_data(altitude, year, n=1000):
    """igrf_data(altitude, year) (magnetic_toolbox.jl:108-121): the n x n x 3
    lat/long map in Tesla.  (The cubic B-spline wrapper of :124-125 is not part of
    the hot path.)"""
    R_E = 6378
    lat = np.linspace(-np.pi / 2, np.pi / 2, n)
    lon = np.linspace(-np.pi, np.pi, n)
    LA, LO = np.meshgrid(lat, lon, indexing="ij")
    r = np.full(LA.size, (altitude + R_E) * 1000.0)
    Bn, Be, Bd = default_engine().igrf12_batch(year, r, LA.ravel(), LO.ravel())
    return np.stack([Bn, Be, Bd], axis=-1).reshape(n, n, 3) / 1.0e9
