"""float64 variants of the parity-path operations (the reference with `jax_enable_x64`, BASELINE.json north_star:
pair sets bit-exact, energies and forces within rel 1e-10).  Functional API over float64 CUDA tensors; every call
goes through a `chx_*_f64` entry point of include/chiron_b200.h.  The throughput engine and the class mirror stay
fp32 like the reference's default configuration.

    displacement / wrap                chiron/neighbors.py:45-112, 116-175
    build_neighborlist / calculate / check   chiron/neighbors.py:595-626, 671-729, 773-787, 864-907
    lj_energy_force                    chiron/potential.py:193-300
    baoab_update / kick                chiron/integrators.py:181-189, 195 (noise handed in)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _dev64(a, device=None):
    if isinstance(a, torch.Tensor):
        dev = a.device if a.is_cuda else (device or _lib.default_device())
        return a.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(device or _lib.default_device())


def _box3(box):
    if box is None:
        return 1.0, 1.0, 1.0
    b = box.detach().cpu().numpy() if isinstance(box, torch.Tensor) else np.asarray(box)
    b = np.asarray(b, dtype=np.float64)
    return (float(b[0, 0]), float(b[1, 1]), float(b[2, 2])) if b.shape == (3, 3) else tuple(float(v) for v in b.reshape(3))


def displacement(x1, x2, box, periodic=True):
    x1, x2 = _dev64(x1).reshape(-1, 3), _dev64(x2).reshape(-1, 3)
    r, d = torch.empty_like(x1), torch.empty(x1.shape[0], dtype=torch.float64, device=x1.device)
    _lib.get_context(x1.device).call("chx_displacement_f64", _lib.ptr(x1), _lib.ptr(x2), x1.shape[0], *_box3(box),
                                     int(periodic), _lib.ptr(r), _lib.ptr(d))
    return r, d


def wrap(x, box):
    x = _dev64(x).reshape(-1, 3)
    out = torch.empty_like(x)
    _lib.get_context(x.device).call("chx_wrap_f64", _lib.ptr(x), x.shape[0], *_box3(box), _lib.ptr(out))
    return out


def build_neighborlist(x, box, cutoff, skin, n_max_neighbors, periodic=True):
    """NeighborListNsqrd.build incl. the growth loop (`>=` instead of the reference's `==`).
    Returns dict(neighbor_list uint32->int64 view, neighbor_mask, n_neighbors, n_max_neighbors)."""
    x = _dev64(x).reshape(-1, 3)
    n, M = x.shape[0], int(n_max_neighbors)
    ctx = _lib.get_context(x.device)
    c = float(cutoff) + float(skin)
    while True:
        nl = torch.empty((n, M), dtype=torch.int32, device=x.device)
        mask = torch.empty((n, M), dtype=torch.int32, device=x.device)
        nn = torch.empty((n,), dtype=torch.int32, device=x.device)
        mx, eq = C.c_int(0), C.c_int(0)
        ctx.call("chx_nlist_build_nsq_f64", _lib.ptr(x), n, *_box3(box), int(periodic), c, M, _lib.ptr(nl), _lib.ptr(mask),
                 _lib.ptr(nn), C.byref(mx), C.byref(eq))
        if mx.value < M:
            break
        M = mx.value + 10
    return dict(neighbor_list=nl, neighbor_mask=mask, n_neighbors=nn, n_max_neighbors=M)


def calculate(x, box, cutoff, neighbor_list, neighbor_mask, periodic=True):
    """NeighborListNsqrd.calculate: (n_neighbors, neighbor_list, mask, dist, r_ij) in float64."""
    x = _dev64(x).reshape(-1, 3)
    n, M = neighbor_list.shape
    n_out = torch.empty((n,), dtype=torch.int32, device=x.device)
    mask = torch.empty((n, M), dtype=torch.int32, device=x.device)
    dist = torch.empty((n, M), dtype=torch.float64, device=x.device)
    rij = torch.empty((n, M, 3), dtype=torch.float64, device=x.device)
    _lib.get_context(x.device).call("chx_nlist_calculate_f64", _lib.ptr(x), n, *_box3(box), int(periodic), float(cutoff), M,
                                    _lib.ptr(neighbor_list), _lib.ptr(neighbor_mask), _lib.ptr(n_out), _lib.ptr(mask),
                                    _lib.ptr(dist), _lib.ptr(rij))
    return n_out, neighbor_list, mask, dist, rij


def check(x, ref_x, box, skin, periodic=True):
    x, ref_x = _dev64(x).reshape(-1, 3), _dev64(ref_x).reshape(-1, 3)
    flag = torch.zeros((), dtype=torch.int32, device=x.device)
    _lib.get_context(x.device).call("chx_nlist_check_f64", _lib.ptr(x), _lib.ptr(ref_x), x.shape[0], *_box3(box),
                                    int(periodic), float(skin) / 2.0, _lib.ptr(flag))
    return bool(flag.item())


def lj_energy_force(x, box, sigma, epsilon, cutoff, neighbor_list, neighbor_mask, periodic=True, want_force=True):
    """LJPotential.compute_energy / compute_force over a NeighborListNsqrd in float64: (energy 0-d, force (N,3))."""
    x = _dev64(x).reshape(-1, 3)
    n, M = neighbor_list.shape
    e = torch.zeros((), dtype=torch.float64, device=x.device)
    F = torch.zeros((n, 3), dtype=torch.float64, device=x.device) if want_force else None
    _lib.get_context(x.device).call("chx_lj_nlist_energy_force_f64", _lib.ptr(x), n, *_box3(box), int(periodic), float(sigma),
                                    float(epsilon), float(cutoff), M, _lib.ptr(neighbor_list), _lib.ptr(neighbor_mask),
                                    _lib.ptr(e), _lib.ptr(F))
    return e, F


def baoab_update(x, v, F, mass, noise, half_dt, a, b, kT, box=None):
    """In-place B-A-O-A of one Langevin step (+ wrap when a box is given) with the (N,3) noise handed in."""
    n = x.shape[0]
    _lib.get_context(x.device).call("chx_baoab_update_f64", _lib.ptr(x), _lib.ptr(v), _lib.ptr(F), _lib.ptr(mass),
                                    _lib.ptr(noise), n, float(half_dt), float(a), float(b), float(kT), *_box3(box),
                                    int(box is not None))


def kick(v, F, mass, half_dt):
    _lib.get_context(v.device).call("chx_kick_f64", _lib.ptr(v), _lib.ptr(F), _lib.ptr(mass), v.shape[0], float(half_dt))
