"""The C restatement (oracle/c, the CPU baseline of bench.py) against the NumPy oracle and the
reference's golden vectors.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import cport, dynamics as dyn, jax_random as jr, pairs, potentials as pot

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_goldens.json")


@pytest.fixture(scope="module")
def goldens():
    with open(GOLD) as fh:
        return json.load(fh)["entries"]


def lattice(n_side, rho_star=0.8, sigma=0.34, seed=0, jitter=0.02):
    n = n_side ** 3
    L = (n / (rho_star / sigma ** 3)) ** (1.0 / 3.0)
    g = (np.arange(n_side) + 0.5) * (L / n_side)
    x = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    x = x + np.random.default_rng(seed).normal(0, jitter, x.shape)
    return x.astype(np.float32), (np.eye(3) * L).astype(np.float32)


def test_threefry_and_normal_bit_exact():
    for seed, n in [(0, 4), (1234, 7), (42, 3000), (7, 12289)]:
        key = jr.PRNGKey(seed)
        assert np.array_equal(cport.random_bits(key, n), jr.random_bits(key, n))
        assert np.array_equal(cport.split(key), jr.split(key))
        a, b = cport.normal(key, (n,)), jr.normal(key, (n,))
        # log1pf (glibc) vs np.log1p (SIMD) may differ in the last place
        assert np.allclose(a, b, rtol=3e-6, atol=1e-7)
    assert np.isclose(cport.normal(jr.PRNGKey(0), (1,))[0], -0.20584226, rtol=1e-6)
    assert np.isclose(cport.normal(jr.PRNGKey(42), (1,))[0], -0.18471177, rtol=1e-6)


def test_space_matches_numpy_and_goldens(goldens):
    g = goldens["space_periodic"]
    box = np.eye(3, dtype=np.float32) * g["box"]
    r, d = cport.displacement(g["p1"], g["p2"], box)
    assert np.array_equal(r, np.asarray(g["r_ij"], np.float32)) and np.array_equal(d, np.asarray(g["dist"], np.float32))
    assert np.array_equal(cport.wrap(g["wrap_in"], box), np.asarray(g["wrap_out"], np.float32))
    rng = np.random.default_rng(3)
    a = rng.uniform(-25, 25, (20000, 3)).astype(np.float32)
    b = rng.uniform(-25, 25, (20000, 3)).astype(np.float32)
    box = np.diag([7.3, 9.1, 11.7]).astype(np.float32)
    r0, d0 = pairs.displacement(a, b, box)
    r1, d1 = cport.displacement(a, b, box)
    assert np.array_equal(r0, r1) and np.array_equal(d0, d1)
    assert np.array_equal(pairs.wrap(a, box), cport.wrap(a, box))


def test_neighborlist_goldens_and_numpy(goldens):
    g = goldens["nlist_cube8"]
    x = np.array([[i, j, k] for i in range(2) for j in range(2) for k in range(2)], np.float32)
    box = np.eye(3, dtype=np.float32) * 10
    out = cport.build_neighborlist(x, box, 2.1, 0.1, 5)
    assert out["n_max_neighbors"] == 17
    assert out["n_neighbors"].tolist() == g["n_neighbors"]
    assert out["neighbor_list"].tolist() == g["neighbor_list_8x17"]
    x, box = lattice(8, seed=5)
    ref = pairs.build_neighborlist(x, box, 1.02, 0.3, 60)
    out = cport.build_neighborlist(x, box, 1.02, 0.3, 60)
    for k in ("neighbor_list", "neighbor_mask", "n_neighbors"):
        assert np.array_equal(ref[k], out[k]), k
    assert cport.check(x + np.float32(0.1), x, box, 0.3) == pairs.check_neighborlist(x + np.float32(0.1), x, box, 0.3)
    assert not cport.check(x, x, box, 0.3)


def test_lj_energy_force_matches_numpy():
    x, box = lattice(8, seed=6)
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    nb = pairs.build_neighborlist(x, box, rc, 0.3, 200)
    e0 = pot.lj_energy_nlist(x, box, sigma, eps, rc, nb["neighbor_list"], nb["neighbor_mask"])
    F0 = pot.lj_force_nlist(x, box, sigma, eps, rc, nb["neighbor_list"], nb["neighbor_mask"])
    e1, F1, n_int = cport.lj_nlist(x, box, sigma, eps, rc, nb["neighbor_list"], nb["neighbor_mask"])
    assert np.isclose(e0, e1, rtol=1e-6)
    assert np.allclose(F0, F1, rtol=1e-5, atol=1e-5 * np.abs(F0).max())
    _, _, nint_ref = pot.lj_energy_force_bruteforce(x, box, sigma, eps, rc, want_force=False)
    assert n_int == nint_ref


def test_langevin_lj_matches_numpy_oracle():
    x, box = lattice(6, seed=8)
    n = x.shape[0]
    sigma, eps, rc, skin = 0.34, 0.238 * 4.184, 1.02, 0.3
    mass = np.full(n, 39.948, np.float32)
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))
    nbr = dyn.OracleNeighborList(box, rc, skin, 100)
    v0 = dyn.maxwell_boltzmann(jr.PRNGKey(5), mass, 300.0)
    key_loop = jr.split(st.key)[1]       # langevin_run draws its loop key with state.new_key()
    xo, vo, key_o, _ = dyn.langevin_run(
        x, v0, mass, 300.0, 0.001, 1.0, st, 10,
        lambda xx: pot.lj_force_nlist(xx, box, sigma, eps, rc, nbr.neighbor_list, nbr.neighbor_mask), nbr=nbr)
    kT = dyn.R_KJ_PER_MOL_K * 300.0
    xc, vc, key_c, stats = cport.langevin_lj(x, v0, mass, box, sigma, eps, rc, skin, 100, kT, 0.001, 1.0, key_loop, 10)
    assert np.array_equal(key_o, key_c)
    assert np.allclose(xo, xc, atol=2e-6) and np.allclose(vo, vc, rtol=1e-4, atol=1e-5)
    assert stats["n_builds"] == nbr.n_builds


def test_cell_grid_accelerator_gives_the_reference_arrays():
    x, box = lattice(12, seed=9)                       # 1728 particles, 4 cells per edge
    x = x + np.float32(30.0) * (np.arange(x.shape[0])[:, None] % 3 == 0)   # some particles outside [0, L)
    cps = np.float32(1.02 + 0.3)
    a = cport.build_rows(x, box, cps, 80)
    b = cport.build_cells(x, box, cps, 80)
    assert a[3] == b[3]
    for u, v in zip(a[:3], b[:3]):
        assert np.array_equal(u, v)
    t, tests = cport.time_reference_build_rows(x, box, cps, stride=7)
    assert tests == sum(x.shape[0] - i - 1 for i in range(0, x.shape[0], 7)) and t >= 0.0
