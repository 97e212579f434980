"""Host logic of the replica-sharded MultiStateSampler: shard bounds, the energy-matrix all-gather and
the swap decisions -- single process and world_size 2 over gloo (CPU).  No GPU needed: the only
library calls are the host-side threefry helpers of the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chiron_b200.multistate import shard_ids, gather_rows, neighbor_swaps, shard_bounds


def test_shard_bounds_cover_all_replicas():
    for R in (1, 3, 7, 64):
        for W in (1, 2, 3, 4, 8):
            seen = []
            for r in range(W):
                lo, hi = shard_bounds(R, W, r)
                assert 0 <= lo <= hi <= R
                seen += list(range(lo, hi))
            assert seen == list(range(R))
            sizes = [shard_bounds(R, W, r)[1] - shard_bounds(R, W, r)[0] for r in range(W)]
            assert max(sizes) - min(sizes) <= 1


def _ladder_matrix(R, rng):
    """u[r, k] = beta_k * U_r for a temperature ladder."""
    U = rng.normal(-1000.0, 30.0, R)
    beta = 1.0 / (0.0083144626 * 300.0 * 2 ** (np.arange(R) / (R - 1.0)))
    return U[:, None] * beta[None, :]


def test_neighbor_swaps_are_permutations_and_deterministic(built_library):
    rng = np.random.default_rng(0)
    R = 16
    u = _ladder_matrix(R, rng)
    states = np.arange(R)
    acc, prop = np.zeros((R, R), np.int64), np.zeros((R, R), np.int64)
    for it in range(1, 40):
        new = neighbor_swaps(u, states, it, seed=7, n_accepted=acc, n_proposed=prop)
        again = neighbor_swaps(u, states, it, seed=7)
        assert np.array_equal(new, again)                       # same key -> same decisions
        assert sorted(new.tolist()) == list(range(R))           # still one replica per state
        moved = np.nonzero(new != states)[0]
        assert np.all(np.abs(new[moved] - states[moved]) == 1)  # neighbour moves only
        states = new
    assert prop.sum() > 0 and 0 < acc.sum() <= prop.sum()
    assert np.array_equal(prop, prop.T) and np.array_equal(acc, acc.T)
    # a swap that lowers the total reduced potential is always accepted
    u2 = np.array([[0.0, 5.0], [5.0, 0.0]])                     # replica 0 prefers state 0 ...
    assert np.array_equal(neighbor_swaps(u2, np.array([1, 0]), 0, seed=1), np.array([0, 1]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, R, out_dir, policy="contiguous"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    u_full = _ladder_matrix(R, rng)
    ids = shard_ids(R, world, rank, policy)
    states = np.arange(R)
    hist = []
    for it in range(1, 6):
        # every rank only knows the rows of its own replicas; U drifts deterministically per sweep
        rows = u_full[ids] * (1.0 + 0.01 * it)
        u = gather_rows(rows, R, policy=policy)
        states = neighbor_swaps(u, states, it, seed=11)
        hist.append(states.copy())
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array(hist))
    np.save(os.path.join(out_dir, f"u{rank}.npy"), u)
    dist.destroy_process_group()


@pytest.mark.parametrize("policy", ["contiguous", "strided"])
def test_two_ranks_gloo_agree_with_single_process(built_library, tmp_path, policy):
    R, world = 7, 2          # uneven shards: 4 + 3
    assert sorted(np.concatenate([shard_ids(R, world, r, policy) for r in range(world)]).tolist()) == list(range(R))
    mp.spawn(_worker, args=(world, _free_port(), R, str(tmp_path), policy), nprocs=world, join=True)
    h0, h1 = np.load(tmp_path / "rank0.npy"), np.load(tmp_path / "rank1.npy")
    assert np.array_equal(h0, h1)                               # identical decisions without a broadcast
    assert np.array_equal(np.load(tmp_path / "u0.npy"), np.load(tmp_path / "u1.npy"))
    # single-process reference
    rng = np.random.default_rng(3)
    u_full = _ladder_matrix(R, rng)
    states = np.arange(R)
    for it in range(1, 6):
        states = neighbor_swaps(u_full * (1.0 + 0.01 * it), states, it, seed=11)
        assert np.array_equal(states, h0[it - 1])
    assert np.allclose(np.load(tmp_path / "u0.npy"), u_full * 1.05)


def _gatherer_worker(rank, world, port, R, K, out_dir, policy):
    import torch
    from chiron_b200.multistate import _RowGatherer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    full = rng.normal(size=(R, K))
    ids = shard_ids(R, world, rank, policy)
    g = _RowGatherer(R, K, policy=policy)
    calls = []
    outs = []
    for it in range(3):            # persistent buffers: several sweeps through the same gatherer
        rows = torch.from_numpy(full[ids] + it)
        outs.append(g(rows, after_enqueue=lambda: calls.append(it)))
    assert calls == [0, 1, 2]      # once per sweep, before the host waits for the matrix
    np.save(os.path.join(out_dir, f"g{rank}.npy"), np.array(outs))
    dist.destroy_process_group()


@pytest.mark.parametrize("policy", ["contiguous", "strided"])
def test_row_gatherer_two_ranks_gloo(built_library, tmp_path, policy):
    """`_RowGatherer` (the device-row all-gather of the REMD exchange, here on host tensors over gloo): uneven shards,
    both sharding policies, the `after_enqueue` hook the table pre-build hangs on."""
    R, K, world = 7, 5, 2
    mp.spawn(_gatherer_worker, args=(world, _free_port(), R, K, str(tmp_path), policy), nprocs=world, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    full = np.random.default_rng(5).normal(size=(R, K))
    for it in range(3):
        assert np.array_equal(g0[it], full + it) and np.array_equal(g1[it], full + it)


def test_row_gatherer_single_process_hook():
    import torch
    from chiron_b200.multistate import _RowGatherer
    rows = torch.arange(12, dtype=torch.float64).reshape(4, 3)
    calls = []
    out = _RowGatherer(4, 3)(rows, after_enqueue=lambda: calls.append(1))
    assert calls == [1] and np.array_equal(out, rows.numpy())


def test_engine_group_rule(monkeypatch):
    """Two engines per GPU from 16 local replicas up (measured, profiles/r02_step_kernel_ncu.md section 7); the
    environment override is clamped to the number of replicas."""
    from chiron_b200.multistate import _engine_group_count
    monkeypatch.delenv("CHX_REMD_ENGINE_GROUPS", raising=False)
    assert [_engine_group_count(n) for n in (1, 8, 15, 16, 32, 64)] == [1, 1, 1, 2, 2, 2]
    monkeypatch.setenv("CHX_REMD_ENGINE_GROUPS", "4")
    assert _engine_group_count(64) == 4 and _engine_group_count(3) == 3
    monkeypatch.setenv("CHX_REMD_ENGINE_GROUPS", "0")
    assert _engine_group_count(8) == 1
