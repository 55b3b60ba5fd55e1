"""ctypes loader for oracle/liboracle_c.so (plain-C restatement, float64 arbiter arithmetic, OpenMP).
TEST INFRASTRUCTURE ONLY — see oracle/oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle_c.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        L = C.CDLL(_PATH)
        L.orc_num_threads.restype = C.c_int
        L.orc_pairwise_sqdist.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_krum_select.restype = C.c_int
        L.orc_krum_select.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_bulyan_select.restype = C.c_int
        L.orc_bulyan_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_bulyan_select_m.restype = C.c_int
        L.orc_bulyan_select_m.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_krum_select_m.restype = C.c_int
        L.orc_krum_select_m.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_trimmed_mean.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_mean.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.orc_alie.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def threads():
    return lib().orc_num_threads()


def _mat(G):
    G = np.asarray(G)
    assert G.dtype == np.float32 and G.ndim == 2 and G.strides[1] == 4
    return G, G.shape[0], G.shape[1], G.strides[0] // 4


def pairwise_sqdist(G):
    G, n, d, ld = _mat(G)
    out = np.empty((n, n), np.float64)
    lib().orc_pairwise_sqdist(G.ctypes.data, n, d, ld, out.ctypes.data)
    return out


def pairwise_dist_as_f32(G):
    """The values the reference stores (np.float32 norms), from float64 sums."""
    return np.sqrt(pairwise_sqdist(G)).astype(np.float32)


def krum_select(dist, users_count, corrupted_count, with_margin=False):
    t = np.ascontiguousarray(dist, np.float64)
    if not with_margin:
        return lib().orc_krum_select(t.ctypes.data, t.shape[0], None, users_count, corrupted_count, None)
    m = C.c_double(0.0)
    idx = lib().orc_krum_select_m(t.ctypes.data, t.shape[0], None, users_count, corrupted_count, C.addressof(m))
    return idx, m.value


def bulyan_select(dist, n, f, with_margins=False):
    t = np.ascontiguousarray(dist, np.float64)
    sel = np.empty(max(n - 2 * f, 1), np.int32)
    mg = np.empty(max(n - 2 * f, 1), np.float64)
    got = lib().orc_bulyan_select_m(t.ctypes.data, n, f, sel.ctypes.data, mg.ctypes.data)
    return (sel[:got].tolist(), mg[:got].tolist()) if with_margins else sel[:got].tolist()


def trimmed_mean(G, corrupted_count, rows=None):
    G, n_total, d, ld = _mat(G)
    r = None if rows is None else np.ascontiguousarray(rows, np.int32)
    n = n_total if r is None else len(r)
    out = np.empty(d, np.float64)
    lib().orc_trimmed_mean(G.ctypes.data, n_total, d, ld, None if r is None else r.ctypes.data, n, corrupted_count,
                           out.ctypes.data)
    return out


def mean(G):
    G, n, d, ld = _mat(G)
    out = np.empty(d, np.float64)
    lib().orc_mean(G.ctypes.data, n, d, ld, out.ctypes.data)
    return out


def alie(G, z):
    G, f, d, ld = _mat(G)
    mu, sd, cr = np.empty(d), np.empty(d), np.empty(d)
    lib().orc_alie(G.ctypes.data, f, d, ld, float(z), mu.ctypes.data, sd.ctypes.data, cr.ctypes.data)
    return cr, mu, sd


def krum(G, users_count, corrupted_count):
    return krum_select(pairwise_dist_as_f32(G).astype(np.float64), users_count, corrupted_count)


def bulyan(G, n, f):
    sel = bulyan_select(pairwise_dist_as_f32(G).astype(np.float64), n, f)
    return trimmed_mean(G, 2 * f, rows=sel), sel
