"""NumPy restatement of chiron's Space / NeighborListNsqrd / PairListNsqrd arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  All arithmetic is IEEE fp32 (or fp64 when
`dtype=np.float64` is passed, the reference's "user enabled x64" mode) with one rounding per
operation, i.e. exactly what XLA:CPU emits for the jitted reference functions:

    displacement   chiron/neighbors.py:45-83   (periodic)   :116-152 (non periodic)
    wrap           chiron/neighbors.py:85-112               :154-175
    build          chiron/neighbors.py:548-729
    calculate      chiron/neighbors.py:731-826
    check          chiron/neighbors.py:828-907
    pair list      chiron/neighbors.py:1018-1289
"""
import numpy as np


def _box_lengths(box, dtype):
    box = np.asarray(box, dtype=dtype)
    if box.shape == (3, 3):
        return np.array([box[0, 0], box[1, 1], box[2, 2]], dtype=dtype)
    return box.reshape(3).astype(dtype)


def jnp_mod(t, L):
    """jnp.mod for floats: lax.rem (C fmod, exact) followed by the sign fix-up
    `where(rem != 0 and (rem < 0) != (L < 0), rem + L, rem)`."""
    rem = np.fmod(t, L)
    fix = (rem != 0) & ((rem < 0) != (L < 0))
    return np.where(fix, (rem + L).astype(t.dtype), rem).astype(t.dtype)


def displacement(x1, x2, box, periodic=True, dtype=np.float32):
    """neighbors.py:69-81: r = mod(x1 - x2 + L/2, L) - L/2 ; d = sqrt(sum r^2) (sequential sum)."""
    x1 = np.asarray(x1, dtype=dtype)
    x2 = np.asarray(x2, dtype=dtype)
    r = (x1 - x2).astype(dtype)
    if periodic:
        if box is None:
            raise ValueError("box_vectors must be provided for a periodic system")
        L = _box_lengths(box, dtype)
        h = (L * dtype(0.5)).astype(dtype)
        r = (jnp_mod((r + h).astype(dtype), L) - h).astype(dtype)
    sq = (r * r).astype(dtype)
    s = ((sq[..., 0] + sq[..., 1]).astype(dtype) + sq[..., 2]).astype(dtype)
    return r, np.sqrt(s).astype(dtype)


def wrap(x, box, periodic=True, dtype=np.float32):
    """neighbors.py:110: x - floor(x / L) * L."""
    x = np.asarray(x, dtype=dtype)
    if not periodic:
        return x
    L = _box_lengths(box, dtype)
    return (x - (np.floor((x / L).astype(dtype)) * L).astype(dtype)).astype(dtype)


def neighbor_rows(x, box, cutoff_plus_skin, periodic=True, dtype=np.float32, rows=None, chunk=2048):
    """Per-row ascending neighbor ids j>i with d_ij < cutoff+skin (neighbors.py:595-602).
    Returns a list of int64 arrays (one per requested row).  Chunked: never materialises N x N."""
    x = np.asarray(x, dtype=dtype)
    n = x.shape[0]
    c = dtype(cutoff_plus_skin)
    rows = np.arange(n) if rows is None else np.asarray(rows)
    out = []
    for s in range(0, rows.size, chunk):
        ii = rows[s:s + chunk]
        _, d = displacement(x[ii][:, None, :], x[None, :, :], box, periodic, dtype)
        m = (d < c) & (ii[:, None] < np.arange(n)[None, :])
        for k in range(ii.size):
            out.append(np.nonzero(m[k])[0])
    return out


def build_neighborlist(x, box, cutoff, skin, n_max_neighbors, periodic=True, dtype=np.float32,
                       grow=True):
    """NeighborListNsqrd.build (neighbors.py:628-729).

    Returns dict(neighbor_list (N,M) u32, neighbor_mask (N,M) i32, n_neighbors (N,) i32,
    n_max_neighbors M).  The growth loop replicates the reference: `while any(n == M): M = max(n)+10`.
    `cutoff + skin` is summed in Python floats first (neighbors.py:674) and rounded once.
    """
    x = np.asarray(x, dtype=dtype)
    n = x.shape[0]
    rows = neighbor_rows(x, box, float(cutoff) + float(skin), periodic, dtype)
    counts = np.array([r.size for r in rows], dtype=np.int32)
    M = int(n_max_neighbors)
    if grow:
        while np.any(counts == M):
            M = int(counts.max()) + 10
    nl = np.zeros((n, M), dtype=np.uint32)
    for i, r in enumerate(rows):
        fill = int(r[0]) if r.size else 0          # argmax(mask): first True, 0 if none
        if fill == i:
            fill += 1                              # neighbors.py:609
        k = min(r.size, M)
        nl[i, :k] = r[:k]
        nl[i, k:] = fill
    mask = (np.arange(M)[None, :] < counts[:, None]).astype(np.int32)
    return dict(neighbor_list=nl, neighbor_mask=mask, n_neighbors=counts, n_max_neighbors=M)


def calculate_neighborlist(x, box, cutoff, neighbor_list, neighbor_mask, periodic=True,
                           dtype=np.float32):
    """NeighborListNsqrd.calculate (neighbors.py:773-826) -> (n, list, mask, dist, r_ij)."""
    x = np.asarray(x, dtype=dtype)
    nl = np.asarray(neighbor_list).astype(np.int64)
    r, d = displacement(x[:, None, :], x[nl], box, periodic, dtype)
    mask = ((d < dtype(cutoff)) & (np.asarray(neighbor_mask) != 0)).astype(np.int32)
    return mask.sum(axis=1).astype(np.int32), neighbor_list, mask, d, r


def check_neighborlist(x, ref_x, box, skin, periodic=True, dtype=np.float32):
    """NeighborListNsqrd.check (neighbors.py:864-907): any ||minimg(x - ref)|| >= skin/2."""
    x = np.asarray(x, dtype=dtype)
    ref_x = np.asarray(ref_x, dtype=dtype)
    if x.shape[0] != ref_x.shape[0]:
        return True
    _, d = displacement(x, ref_x, box, periodic, dtype)
    return bool(np.any(d >= dtype(float(skin) / 2.0)))


def build_pairlist(n):
    """PairListNsqrd.build (neighbors.py:1018-1104): all_pairs (N,N-1) = j != i ascending,
    reduction_mask = i < j."""
    ids = np.arange(n, dtype=np.uint32)
    jj = np.broadcast_to(ids, (n, n))
    all_pairs = jj[~np.eye(n, dtype=bool)].reshape(n, n - 1).astype(np.uint32)
    reduction_mask = ids[:, None] < all_pairs
    return all_pairs, reduction_mask


def calculate_pairlist(x, box, cutoff, all_pairs, reduction_mask, periodic=True, dtype=np.float32):
    """PairListNsqrd.calculate (neighbors.py:1106-1269); cutoff=None -> mask = reduction mask."""
    x = np.asarray(x, dtype=dtype)
    r, d = displacement(x[:, None, :], x[all_pairs.astype(np.int64)], box, periodic, dtype)
    if cutoff is None:
        mask = reduction_mask.astype(np.int32)
    else:
        mask = ((d < dtype(cutoff)) & reduction_mask).astype(np.int32)
    return mask.sum(axis=1).astype(np.int32), all_pairs, mask, d, r


def pair_set(neighbor_list, n_neighbors):
    """The list as a set of (i, j) pairs -- the object the bit-exactness contract is stated on."""
    out = set()
    for i, k in enumerate(np.asarray(n_neighbors)):
        for j in np.asarray(neighbor_list)[i, :int(k)]:
            out.add((i, int(j)))
    return out


def pair_keys(neighbor_list, n_neighbors):
    """Sorted int64 keys i*N+j of all listed pairs (vectorised pair_set for big systems)."""
    nl = np.asarray(neighbor_list).astype(np.int64)
    nn = np.asarray(n_neighbors).astype(np.int64)
    n, M = nl.shape
    valid = np.arange(M)[None, :] < nn[:, None]
    keys = (np.arange(n, dtype=np.int64)[:, None] * n + nl)[valid]
    return np.sort(keys)
