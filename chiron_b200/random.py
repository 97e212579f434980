"""The handful of `jax.random` calls chiron makes, on the reference's legacy-threefry stream.

Keys are `(2,) uint32` NumPy arrays in JAX's raw key format, so seeds, key threading and every
random number are interchangeable with the reference (`chiron/utils.py:29-38`,
`chiron/states.py:150-154`).  Bulk draws run on the GPU (`chx_random_normal/uniform`); scalar
draws used for host-side decisions (`chiron/mcmc.py:544,967`) are computed on the host from
`chx_random_bits_host` with the same fp32 formula.
"""
import numpy as np
import torch

from . import _lib


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def _as_key(key) -> np.ndarray:
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().numpy()
    key = np.asarray(key)
    if key.shape != (2,):
        raise ValueError(f"PRNG key must have shape (2,), got {key.shape}")
    return key.astype(np.uint32)


def split(key, num: int = 2):
    if num != 2:
        bits = _lib.random_bits_host(_as_key(key), 2 * num)
        return bits.reshape(num, 2)
    carried, sub = _lib.split_host(_as_key(key))
    return np.stack([carried, sub])


def split_many(keys):
    """`random.split(key)` for n keys at once: (carried (n,2), subkeys (n,2)) -- one library call."""
    keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint32).reshape(-1, 2))
    out = _lib.split_host_n(keys)
    return out[:, 0:2].copy(), out[:, 2:4].copy()


def _numel(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def normal(key, shape=(), device=None) -> torch.Tensor:
    key = _as_key(key)
    shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
    ctx = _lib.get_context(device)
    out = torch.empty(shape, dtype=torch.float32, device=ctx.device)
    ctx.call("chx_random_normal", int(key[0]), int(key[1]), _numel(shape), _lib.ptr(out))
    return out


def uniform_host_n(key, n, minval=0.0, maxval=1.0) -> np.ndarray:
    """`random.uniform(key, (n,), minval, maxval)` evaluated on the host in fp32."""
    bits = _lib.random_bits_host(_as_key(key), int(n))
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, (f * np.float32(hi - lo)).astype(np.float32) + lo).astype(np.float32)


def uniform_host(key, minval=0.0, maxval=1.0) -> np.float32:
    """Scalar `random.uniform(key, minval=, maxval=)` evaluated on the host in fp32."""
    return np.float32(uniform_host_n(key, 1, minval, maxval)[0])


def uniform(key, shape=(), minval=0.0, maxval=1.0, device=None):
    shape = tuple(int(s) for s in (shape if hasattr(shape, "__len__") else (shape,)))
    if shape == ():
        return uniform_host(key, minval, maxval)
    key = _as_key(key)
    ctx = _lib.get_context(device)
    out = torch.empty(shape, dtype=torch.float32, device=ctx.device)
    ctx.call("chx_random_uniform", int(key[0]), int(key[1]), _numel(shape), float(minval),
             float(maxval), _lib.ptr(out))
    return out
