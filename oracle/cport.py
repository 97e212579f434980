"""ctypes binding of oracle/c/chiron_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The C restatement is the multi-threaded CPU baseline of bench.py (`cpu_baseline`, `--impl reference`)
and is itself checked against the NumPy oracle in tests/test_oracle_c.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libchiron_oracle.so")
_lib = None

_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class LangevinStats(C.Structure):
    _fields_ = [("n_builds", C.c_int), ("M", C.c_int), ("t_build_s", C.c_double), ("t_steps_s", C.c_double),
                ("energy", C.c_double), ("p_cand", C.c_longlong), ("p_int", C.c_longlong)]


def build():
    src = os.path.join(_HERE, "c", "chiron_oracle.c")
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "c")], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_displacement.argtypes = [_fp, _fp, C.c_long, _fp, C.c_int, _fp, _fp]
        L.orc_wrap.argtypes = [_fp, C.c_long, _fp]
        L.orc_nlist_build_rows.argtypes = [_fp, C.c_int, _fp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, _ip]
        L.orc_nlist_build_rows.restype = C.c_int
        L.orc_nlist_build_cells.argtypes = [_fp, C.c_int, _fp, C.c_float, C.c_int, C.c_void_p, C.c_void_p, _ip]
        L.orc_nlist_build_cells.restype = C.c_int
        L.orc_nlist_time_rows.argtypes = [_fp, C.c_int, _fp, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        L.orc_nlist_time_rows.restype = C.c_double
        L.orc_set_build_mode.argtypes = [C.c_int]
        L.orc_nlist_check.argtypes = [_fp, _fp, C.c_int, _fp, C.c_int, C.c_float]
        L.orc_lj_nlist.argtypes = [_fp, C.c_int, _fp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                   C.c_int, _up, _ip, C.c_void_p, C.POINTER(C.c_longlong)]
        L.orc_lj_nlist.restype = C.c_double
        L.orc_random_bits.argtypes = [_up, C.c_long, _up]
        L.orc_split.argtypes = [_up, _up]
        L.orc_normal.argtypes = [_up, C.c_long, _fp]
        L.orc_langevin_lj.argtypes = [_fp, _fp, _fp, C.c_int, _fp, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_int, C.c_float, C.c_float, C.c_float, _up, C.c_int,
                                      C.POINTER(LangevinStats)]
        _lib = L
    return _lib


def _box3(box):
    b = np.asarray(box, dtype=np.float32)
    return np.ascontiguousarray(np.diag(b) if b.shape == (3, 3) else b.reshape(3))


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def num_threads():
    return int(lib().orc_num_threads())


def use_all_cores():
    """Use every core this process may run on (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(n)
    return num_threads()


def displacement(x1, x2, box, periodic=True):
    x1, x2 = _f(x1).reshape(-1, 3), _f(x2).reshape(-1, 3)
    r = np.empty_like(x1)
    d = np.empty(x1.shape[0], np.float32)
    lib().orc_displacement(x1, x2, x1.shape[0], _box3(box) if periodic else np.ones(3, np.float32), int(periodic), r, d)
    return r, d


def wrap(x, box):
    x = _f(x).copy()
    lib().orc_wrap(x, x.shape[0], _box3(box))
    return x


def build_rows(x, box, cutoff_plus_skin, M, row0=0, row1=None, periodic=True, want_list=True):
    x = _f(x)
    n = x.shape[0]
    row1 = n if row1 is None else row1
    nr = row1 - row0
    nn = np.empty(nr, np.int32)
    if want_list:
        nl = np.empty((nr, M), np.uint32)
        mask = np.empty((nr, M), np.int32)
        pl, pm = nl.ctypes.data, mask.ctypes.data
    else:
        nl = mask = None
        pl = pm = None
    b = _box3(box) if periodic else np.ones(3, np.float32)
    mx = lib().orc_nlist_build_rows(x, n, b, int(periodic), float(cutoff_plus_skin), int(M), row0, row1, pl, pm, nn)
    return nl, mask, nn, int(mx)


def build_neighborlist(x, box, cutoff, skin, n_max_neighbors, periodic=True):
    """NeighborListNsqrd.build incl. the growth loop; same dict as oracle.pairs.build_neighborlist."""
    M = int(n_max_neighbors)
    cps = np.float32(float(cutoff) + float(skin))
    while True:
        nl, mask, nn, mx = build_rows(x, box, cps, M, periodic=periodic)
        if not np.any(nn == M):
            break
        M = mx + 10
    return dict(neighbor_list=nl, neighbor_mask=mask, n_neighbors=nn, n_max_neighbors=M)


def build_cells(x, box, cutoff_plus_skin, M):
    """Cell-grid builder (set-up accelerator, same arrays as the reference build)."""
    x = _f(x)
    n = x.shape[0]
    nl = np.empty((n, M), np.uint32)
    mask = np.empty((n, M), np.int32)
    nn = np.empty(n, np.int32)
    mx = lib().orc_nlist_build_cells(x, n, _box3(box), float(cutoff_plus_skin), int(M), nl.ctypes.data, mask.ctypes.data, nn)
    return nl, mask, nn, int(mx)


def time_reference_build_rows(x, box, cutoff_plus_skin, stride, row0=0):
    """(seconds, pair tests) of the reference's O(N^2) row loop on every `stride`-th row."""
    x = _f(x)
    tests = C.c_longlong(0)
    t = lib().orc_nlist_time_rows(x, x.shape[0], _box3(box), float(cutoff_plus_skin), int(row0), int(stride), C.byref(tests))
    return float(t), int(tests.value)


def set_build_mode(mode):
    """0: langevin_lj rebuilds with the reference's O(N^2) loop; 1: with the cell-grid accelerator."""
    lib().orc_set_build_mode(int(mode))


def check(x, ref, box, skin, periodic=True):
    x, ref = _f(x), _f(ref)
    b = _box3(box) if periodic else np.ones(3, np.float32)
    return bool(lib().orc_nlist_check(x, ref, x.shape[0], b, int(periodic), np.float32(float(skin) / 2.0)))


def lj_nlist(x, box, sigma, epsilon, cutoff, neighbor_list, neighbor_mask, row0=0, periodic=True, want_force=True):
    x = _f(x)
    n = x.shape[0]
    nl = np.ascontiguousarray(neighbor_list, dtype=np.uint32)
    mask = np.ascontiguousarray(neighbor_mask, dtype=np.int32)
    F = np.zeros((n, 3), np.float32) if want_force else None
    nint = C.c_longlong(0)
    b = _box3(box) if periodic else np.ones(3, np.float32)
    e = lib().orc_lj_nlist(x, n, b, int(periodic), sigma, epsilon, cutoff, nl.shape[1], row0, row0 + nl.shape[0],
                           nl, mask, F.ctypes.data if want_force else None, C.byref(nint))
    return float(e), F, int(nint.value)


def random_bits(key, n):
    out = np.empty(max(n, 1), np.uint32)
    lib().orc_random_bits(np.ascontiguousarray(key, dtype=np.uint32), n, out)
    return out[:n]


def split(key):
    out = np.empty(4, np.uint32)
    lib().orc_split(np.ascontiguousarray(key, dtype=np.uint32), out)
    return out.reshape(2, 2)


def normal(key, shape):
    n = int(np.prod(shape))
    out = np.empty(max(n, 1), np.float32)
    lib().orc_normal(np.ascontiguousarray(key, dtype=np.uint32), n, out)
    return out[:n].reshape(shape)


def langevin_lj(x, v, mass, box, sigma, epsilon, cutoff, skin, n_max_neighbors, kT, dt, gamma, key, nsteps):
    """LangevinIntegrator.run body (loop key in, loop key out).  Returns (x, v, key, stats dict)."""
    x, v, mass = _f(x).copy(), _f(v).copy(), _f(mass)
    key = np.ascontiguousarray(key, dtype=np.uint32).copy()
    st = LangevinStats()
    lib().orc_langevin_lj(x, v, mass, x.shape[0], _box3(box), sigma, epsilon, cutoff, skin, int(n_max_neighbors),
                          kT, dt, gamma, key, int(nsteps), C.byref(st))
    return x, v, key, {f: getattr(st, f) for f, _ in LangevinStats._fields_}
