"""NumPy restatement of the `jax.random` calls chiron makes (legacy, non-partitionable threefry).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Third-party arithmetic: `jax` is an un-pinned dependency of the reference
(`devtools/conda-envs/test_env.yaml:10`).  The reference's golden vectors
(`chiron/tests/test_mcmc.py:81-84`, `:451-452`) pin it to the legacy stream
(`jax_threefry_partitionable=False`), which is what is restated here from the published
algorithm (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11, Threefry-2x32
with 20 rounds; JAX's `_threefry_split`, `_threefry_random_bits_original`, `_uniform`, `_normal_real`
and XLA's single-precision `ErfInv` polynomial after M. Giles).

Call sites in the reference:
    random.PRNGKey / split     chiron/utils.py:29-38, chiron/states.py:150-154, chiron/integrators.py:179
    random.normal              chiron/integrators.py:185, chiron/utils.py:142, chiron/mcmc.py:734
    random.uniform             chiron/mcmc.py:544, chiron/mcmc.py:967
"""
import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def _rotl(x, r):
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(key, x0, x1):
    """Threefry-2x32, 20 rounds.  key: (2,) u32; x0, x1: u32 arrays of equal shape."""
    with np.errstate(over="ignore"):
        k0, k1 = _U32(key[0]), _U32(key[1])
        ks = (k0, k1, _U32(k0 ^ k1 ^ _U32(0x1BD11BDA)))
        x0 = (np.asarray(x0, dtype=_U32) + ks[0]).astype(_U32)
        x1 = (np.asarray(x1, dtype=_U32) + ks[1]).astype(_U32)
        for g in range(1, 6):
            for r in _ROT[(g - 1) % 2]:
                x0 = (x0 + x1).astype(_U32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[g % 3]).astype(_U32)
            x1 = (x1 + ks[(g + 1) % 3] + _U32(g)).astype(_U32)
    return x0, x1


def random_bits(key, n: int) -> np.ndarray:
    """`_threefry_random_bits_original` for 32-bit output: counters 0..n-1, odd n padded with one 0,
    first half -> x0 lanes, second half -> x1 lanes, outputs concatenated and truncated to n."""
    n = int(n)
    if n == 0:
        return np.zeros(0, dtype=_U32)
    counts = np.arange(n, dtype=_U32)
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, dtype=_U32)])
    half = counts.size // 2
    o0, o1 = threefry2x32(key, counts[:half], counts[half:])
    return np.concatenate([o0, o1])[:n]


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split: (num, 2) u32.  chiron keeps row 0 and hands out row 1."""
    return random_bits(key, 2 * num).reshape(num, 2)


def _bits_to_unit_float(bits):
    f = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32)
    return f - np.float32(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    shape = tuple(np.atleast_1d(shape).astype(int)) if shape != () else ()
    n = int(np.prod(shape)) if shape else 1
    f = _bits_to_unit_float(random_bits(key, n))
    lo, hi = np.float32(minval), np.float32(maxval)
    out = np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)
    return out.reshape(shape) if shape else out.reshape(())


_ERFINV_SMALL = np.array(
    [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
     -0.00125372503, -0.00417768164, 0.246640727, 1.50140941], dtype=np.float32)
_ERFINV_LARGE = np.array(
    [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
     -0.0076224613, 0.00943887047, 1.00167406, 2.83297682], dtype=np.float32)


def erfinv_f32(x) -> np.ndarray:
    """XLA's fp32 ErfInv (Giles' single-precision polynomial)."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-(x * x)).astype(np.float32)
        small = w < np.float32(5.0)
        ws = w - np.float32(2.5)
        wl = np.sqrt(w).astype(np.float32) - np.float32(3.0)
        ww = np.where(small, ws, wl).astype(np.float32)
        p = np.where(small, _ERFINV_SMALL[0], _ERFINV_LARGE[0]).astype(np.float32)
        for k in range(1, 9):
            c = np.where(small, _ERFINV_SMALL[k], _ERFINV_LARGE[k]).astype(np.float32)
            p = (c + p * ww).astype(np.float32)
        out = (p * x).astype(np.float32)
        out = np.where(np.abs(x) == np.float32(1.0), np.float32(np.inf) * x, out)
    return out.astype(np.float32)


def normal(key, shape=()) -> np.ndarray:
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(key, shape, minval=lo, maxval=1.0)
    return (np.float32(np.sqrt(2.0)) * erfinv_f32(u)).astype(np.float32)
