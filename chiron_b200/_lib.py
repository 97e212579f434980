"""ctypes binding of libchiron_b200.so (see include/chiron_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present every
compute entry point raises.  The library is built in-tree by `chiron_b200.build` (nvcc, sm_100a).
"""
import ctypes as C
import os
import threading

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libchiron_b200.so")

_lib = None
_lock = threading.Lock()
_contexts = {}

c_float_p = C.POINTER(C.c_float)
_P = C.c_void_p
_I = C.c_int
_L = C.c_longlong
_F = C.c_float
_U = C.c_uint32
_D = C.c_double


class ChironB200Error(RuntimeError):
    pass


# name -> argtypes; every symbol declared in include/chiron_b200.h must appear here
SIGNATURES = {
    "chx_version": [],
    "chx_context_create": [_I, _P, C.POINTER(_P)],
    "chx_context_set_stream": [_P, _P],
    "chx_context_destroy": [_P],
    "chx_synchronize": [_P],
    "chx_launch_count": [_P],
    "chx_displacement": [_P, _P, _P, _L, _F, _F, _F, _I, _P, _P],
    "chx_wrap": [_P, _P, _L, _F, _F, _F, _I, _P],
    "chx_nlist_build_nsq": [_P, _P, _I, _F, _F, _F, _I, _F, _I, _P, _P, _P, C.POINTER(_I), C.POINTER(_I)],
    "chx_nlist_build_cell": [_P, _P, _I, _F, _F, _F, _I, _F, _I, _P, _P, _P, C.POINTER(_I), C.POINTER(_I)],
    "chx_nlist_calculate": [_P, _P, _I, _F, _F, _F, _I, _F, _I, _P, _P, _P, _P, _P, _P],
    "chx_nlist_check": [_P, _P, _P, _I, _F, _F, _F, _I, _F, _P],
    "chx_pairlist_build": [_P, _I, _P, _P],
    "chx_pairlist_calculate": [_P, _P, _I, _F, _F, _F, _I, _F, _P, _P, _P, _P],
    "chx_lj_nlist_energy_force": [_P, _P, _I, _F, _F, _F, _I, _P, _P, _I, _F, _F, _F, _P, _P],
    "chx_lj_allpairs_energy_force": [_P, _P, _I, _F, _F, _F, _I, _F, _F, _F, _P, _P],
    "chx_ho_energy_force": [_P, _P, _I, _P, _I, _F, _F, _P, _P],
    "chx_lj_subset_delta_energy": [_P, _P, _P, _I, _P, _I, _F, _F, _F, _I, _F, _F, _F, _P],
    "chx_threefry_split_host": [C.POINTER(_U), C.POINTER(_U)],
    "chx_random_bits_host": [C.POINTER(_U), _L, C.POINTER(_U)],
    "chx_threefry_split_host_n": [C.POINTER(_U), _I, C.POINTER(_U)],
    "chx_random_normal": [_P, _U, _U, _L, _P],
    "chx_random_uniform": [_P, _U, _U, _L, _F, _F, _P],
    "chx_baoab_update": [_P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _U, _U, _I, _F, _F, _F, _I, _P, _F, _P],
    "chx_kick": [_P, _P, _P, _P, _I, _F],
    "chx_init_velocities": [_P, _P, _P, _I, _F, _U, _U],
    "chx_mc_displace": [_P, _P, _I, _U, _U, _F, _P, _F, _F, _F, _I, _P],
    "chx_scale": [_P, _P, _L, _F, _P],
    "chx_mc_displace_run": [_P, _P, _P, _P, _P, _P, _I],
    "chx_mc_barostat_run": [_P, _P, _P, _P, _P, _P, _I],
    "chx_lj_nlist_energy_force_mixed": [_P, _P, _I, _F, _F, _F, _I, _P, _P, _I, _P, _P, _F, _I, _F, _P, _P],
    # x64 variants
    "chx_displacement_f64": [_P, _P, _P, _L, _D, _D, _D, _I, _P, _P],
    "chx_wrap_f64": [_P, _P, _L, _D, _D, _D, _P],
    "chx_nlist_build_nsq_f64": [_P, _P, _I, _D, _D, _D, _I, _D, _I, _P, _P, _P, C.POINTER(_I), C.POINTER(_I)],
    "chx_nlist_calculate_f64": [_P, _P, _I, _D, _D, _D, _I, _D, _I, _P, _P, _P, _P, _P, _P],
    "chx_nlist_check_f64": [_P, _P, _P, _I, _D, _D, _D, _I, _D, _P],
    "chx_lj_nlist_energy_force_f64": [_P, _P, _I, _D, _D, _D, _I, _D, _D, _D, _I, _P, _P, _P, _P],
    "chx_baoab_update_f64": [_P, _P, _P, _P, _P, _P, _I, _D, _D, _D, _D, _D, _D, _D, _I],
    "chx_kick_f64": [_P, _P, _P, _P, _I, _D],
}


class McDisplaceArgs(C.Structure):
    """chx_mc_displace_args (include/chiron_b200.h)."""
    _fields_ = [("n", _I), ("potential", _I), ("periodic", _I), ("lx", _F), ("ly", _F), ("lz", _F),
                ("sigma", _F), ("epsilon", _F), ("cutoff", _F), ("subset_mask", _P), ("subset_ids", _P),
                ("n_subset", _I), ("neighbor_list", _P), ("n_neighbors", _P), ("M", _I),
                ("ref_positions", _P), ("skin", _F), ("x0", _P), ("n0", _I), ("k", _F), ("U0", _F),
                ("beta", C.c_double), ("pv", C.c_double)]


class McState(C.Structure):
    """chx_mc_state (include/chiron_b200.h), 64 bytes."""
    _fields_ = [("key", _U * 2), ("sel", C.c_int32), ("have_u", C.c_int32), ("u_current", _F),
                ("sigma_disp", _F), ("n_accepted", C.c_int32), ("n_proposed", C.c_int32),
                ("moves_done", C.c_int32), ("halt", C.c_int32), ("nan_seen", C.c_int32),
                ("reserved", C.c_int32 * 5)]


MC_LJ_NLIST, MC_LJ_ALLPAIRS, MC_HO, MC_IDEAL, MC_LJ_SUBSET_DELTA = range(5)


class McBarostatArgs(C.Structure):
    """chx_mc_barostat_args (include/chiron_b200.h)."""
    _fields_ = [("n", _I), ("sigma", _F), ("epsilon", _F), ("cutoff", _F), ("cutoff_plus_skin", _F), ("M", _I),
                ("neighbor_list", _P * 2), ("neighbor_mask", _P * 2), ("n_neighbors", _P * 2),
                ("beta", C.c_double), ("pressure", C.c_double), ("ncell_capacity", _I),
                ("superset_list", _P), ("superset_nn", _P)]


class McBaroState(C.Structure):
    """chx_mc_baro_state (include/chiron_b200.h), 64 bytes."""
    _fields_ = [("key", _U * 2), ("sel", C.c_int32), ("have_u", C.c_int32), ("u_current", _F),
                ("volume_max_scale", _F), ("n_accepted", C.c_int32), ("n_proposed", C.c_int32),
                ("moves_done", C.c_int32), ("halt", C.c_int32), ("nan_seen", C.c_int32), ("box", _F * 3),
                ("last_volume", _F), ("reserved", C.c_int32)]
_RESTYPES = {"chx_launch_count": _L, "chx_last_error_string": C.c_char_p}


def load_library():
    """dlopen the in-tree shared library and declare prototypes.  Raises if it is missing."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ChironB200Error(
                f"{LIB_PATH} not found: build it with `python -m chiron_b200.build` "
                "(chiron_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.chx_last_error_string.restype = C.c_char_p
        lib.chx_last_error_string.argtypes = []
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, _I)
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        msg = load_library().chx_last_error_string().decode(errors="replace")
        raise ChironB200Error(f"libchiron_b200 error {rc}: {msg}")


class Context:
    """A chx_ctx bound to (device, torch current stream at creation)."""

    def __init__(self, device_index):
        lib = load_library()
        if not torch.cuda.is_available():
            raise ChironB200Error("no CUDA device: chiron_b200 runs on sm_100a only (no CPU fallback)")
        self.lib = lib
        self.device = torch.device("cuda", device_index)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
        h = _P()
        check(lib.chx_context_create(device_index, _P(stream), C.byref(h)))
        self.handle = h
        self._stream = stream

    def sync_stream(self):
        """Follow torch's current stream."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != self._stream:
            check(self.lib.chx_context_set_stream(self.handle, _P(s)))
            self._stream = s

    @property
    def launches(self):
        return int(self.lib.chx_launch_count(self.handle))

    def call(self, name, *args):
        self.sync_stream()
        check(getattr(self.lib, name)(self.handle, *args))


def get_context(device=None) -> Context:
    if device is None:
        if not torch.cuda.is_available():
            raise ChironB200Error("no CUDA device: chiron_b200 runs on sm_100a only (no CPU fallback)")
        idx = torch.cuda.current_device()
    else:
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
    ctx = _contexts.get(idx)
    if ctx is None:
        ctx = Context(idx)
        _contexts[idx] = ctx
    return ctx


def default_device():
    if not torch.cuda.is_available():
        raise ChironB200Error("no CUDA device: chiron_b200 runs on sm_100a only (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return _P(0)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ChironB200Error("expected a CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        raise ChironB200Error(f"expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ChironB200Error("expected a contiguous tensor")
    return _P(t.data_ptr())


def as_device_f32(a, device=None):
    """Anything array-like -> contiguous float32 CUDA tensor (no copy if already one)."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda and a.dtype == torch.float32 and a.is_contiguous():
            return a
        dev = a.device if a.is_cuda else (device or default_device())
        return a.to(device=dev, dtype=torch.float32).contiguous()
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return torch.from_numpy(arr).to(device or default_device())


# ---- host-side jax.random helpers (keys are (2,) uint32 NumPy arrays in JAX's raw format) ----------
def split_host(key):
    lib = load_library()
    k = (_U * 2)(int(key[0]), int(key[1]))
    out = (_U * 4)()
    check(lib.chx_threefry_split_host(k, out))
    return (np.array([out[0], out[1]], dtype=np.uint32), np.array([out[2], out[3]], dtype=np.uint32))


def split_host_n(keys):
    """(n,2) uint32 keys -> (n,4) uint32: [carried key, subkey] of random.split for every row."""
    lib = load_library()
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    out = np.empty((keys.shape[0], 4), dtype=np.uint32)
    check(lib.chx_threefry_split_host_n(keys.ctypes.data_as(C.POINTER(_U)), keys.shape[0],
                                        out.ctypes.data_as(C.POINTER(_U))))
    return out


def random_bits_host(key, n):
    lib = load_library()
    k = (_U * 2)(int(key[0]), int(key[1]))
    out = (_U * max(1, n))()
    check(lib.chx_random_bits_host(k, n, out))
    return np.frombuffer(out, dtype=np.uint32, count=n).copy()
