"""Pins the CPU oracle against every golden vector chiron's own tests hold for the hot path
(tests/golden/reference_goldens.json, extracted from /root/reference by tests/golden/make_golden.py)
and against the published jax.random known answers.  CPU only."""
import numpy as np

from oracle import dynamics as dyn
from oracle import jax_random as jr
from oracle import pairs, potentials as pot

f32 = np.float32
BOX10 = np.eye(3, dtype=f32) * 10


def test_threefry_known_answers():
    # JAX legacy (non-partitionable) stream, SURVEY.md App. A.6
    assert jr.split(jr.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert np.isclose(jr.normal(jr.PRNGKey(0)), -0.20584226, rtol=0, atol=1e-7)
    assert np.isclose(jr.normal(jr.PRNGKey(42)), -0.18471177, rtol=0, atol=1e-7)
    assert np.isclose(jr.uniform(jr.PRNGKey(0)), 0.41845703, rtol=0, atol=1e-7)


def test_random_bits_layout_odd_even():
    key = jr.PRNGKey(7)
    for n in (1, 2, 3, 6, 7, 30, 31):
        bits = jr.random_bits(key, n)
        assert bits.shape == (n,)
        # element e < ceil(n/2) is out0 of block e, otherwise out1 of block e - ceil(n/2)
        half = (n + 1) // 2
        for e in range(n):
            blk = e if e < half else e - half
            c1 = blk + half
            o0, o1 = jr.threefry2x32(key, np.array([blk], np.uint32), np.array([c1 if c1 < n else 0], np.uint32))
            assert bits[e] == (o0[0] if e < half else o1[0])


def test_space_goldens(goldens):
    g = goldens["space_periodic"]
    r, d = pairs.displacement(np.array(g["p1"], f32), np.array(g["p2"], f32), BOX10)
    assert np.array_equal(r, np.array(g["r_ij"], f32)) and np.array_equal(d, np.array(g["dist"], f32))
    w = pairs.wrap(np.array(g["wrap_in"], f32), BOX10)
    assert np.array_equal(w, np.array(g["wrap_out"], f32))
    r, d = pairs.displacement(np.array(g["p1"], f32), np.array(g["p2"], f32), BOX10, periodic=False)
    assert np.array_equal(r, np.array([[-1, 0, 0], [-6, 0, 0]], f32)) and np.array_equal(d, np.array([1, 6], f32))


def test_neighborlist_pair_golden(goldens):
    g = goldens["nlist_pair2"]
    x = np.array([[0, 0, 0], [1, 0, 0]], f32)
    out = pairs.build_neighborlist(x, BOX10, 1.1, 0.1, 5)
    assert out["neighbor_list"].tolist() == g["neighbor_list"]
    assert out["neighbor_mask"].tolist() == g["neighbor_mask"]
    assert out["n_neighbors"].tolist() == g["n_neighbors"]
    n, nl, mask, dist, rij = pairs.calculate_neighborlist(x, BOX10, 1.1, out["neighbor_list"], out["neighbor_mask"])
    assert n.tolist() == [1, 0] and mask.tolist() == g["neighbor_mask"]
    assert np.all(dist == 1.0)
    assert np.all(rij[0] == np.array([-1, 0, 0], f32)) and np.all(rij[1] == np.array([1, 0, 0], f32))
    assert pairs.check_neighborlist(x, x, BOX10, 0.1) is False
    assert pairs.check_neighborlist(x + f32(0.1), x, BOX10, 0.1) is True
    assert pairs.check_neighborlist(np.zeros((3, 3), f32), x, BOX10, 0.1) is True


def cube8():
    g = np.mgrid[0:2, 0:2, 0:2].astype(f32) * f32(2.0) / f32(2)
    return np.stack(g.reshape(3, -1), axis=1).astype(f32)


def test_neighborlist_cube_golden(goldens):
    g = goldens["nlist_cube8"]
    x = cube8()
    out = pairs.build_neighborlist(x, BOX10, 2.1, 0.1, 5)
    assert out["n_neighbors"].tolist() == g["n_neighbors"]
    n, *_ = pairs.calculate_neighborlist(x, BOX10, 2.1, out["neighbor_list"], out["neighbor_mask"])
    assert n.tolist() == g["n_neighbors"]
    out = pairs.build_neighborlist(x, BOX10, 1.1, 1.1, 5)
    assert out["n_max_neighbors"] == 17
    assert out["neighbor_list"].tolist() == g["neighbor_list_8x17"]
    n, *_ = pairs.calculate_neighborlist(x, BOX10, 1.1, out["neighbor_list"], out["neighbor_mask"])
    assert n.tolist() == g["n_interacting_cutoff_1p1"]


def test_pairlist_cube_golden(goldens):
    g = goldens["pairlist_cube8"]
    x = cube8()
    ap, red = pairs.build_pairlist(8)
    assert ap.tolist() == g["all_pairs"]
    n, _, mask, dist, _ = pairs.calculate_pairlist(x, BOX10, 2.1, ap, red)
    assert n.tolist() == [7, 6, 5, 4, 3, 2, 1, 0]
    assert np.allclose(dist, np.array(g["distances"], f32), rtol=0, atol=1e-7)
    ap2, red2 = pairs.build_pairlist(2)
    assert ap2.tolist() == [[1], [0]] and red2.tolist() == [[True], [False]]
    x2 = np.array([[0, 0, 0], [1, 0, 0]], f32)
    assert pairs.calculate_pairlist(x2, BOX10, 0.5, ap2, red2)[2].tolist() == [[0], [0]]
    assert pairs.calculate_pairlist(x2, BOX10, None, ap2, red2)[2].tolist() == [[1], [0]]


def test_ho_energy_golden(goldens):
    g = goldens["ho_energies"]
    k = g["k_kcal_per_mol_A2"] * 4.184 * 100.0
    for p, e in zip(g["positions_A"], g["energies"]):
        x = (np.array([p], dtype=np.float64) * 0.1).astype(f32)
        assert np.isclose(pot.ho_energy(x, np.zeros((1, 3), f32), k, 0.0), e, rtol=1e-5, atol=1e-8)


def test_lj_two_particle_analytic():
    sigma, eps, cutoff, skin = 1.0, 1.0, 3.0, 0.5
    for i in range(1, 11):
        x = np.array([[0, 0, 0], [i * 0.25 * 2 ** (1 / 6), 0, 0]], f32)
        d = np.linalg.norm((x[0] - x[1]).astype(np.float64))
        e_ref = 4 * eps * ((sigma / d) ** 12 - (sigma / d) ** 6) if d < cutoff else 0.0
        f_ref = (24 * eps / d ** 2 * (2 * (sigma / d) ** 12 - (sigma / d) ** 6)) * (x[0] - x[1]) if d < cutoff else 0 * x[0]
        nl = pairs.build_neighborlist(x, BOX10, cutoff, skin, 5)
        e = pot.lj_energy_nlist(x, BOX10, sigma, eps, cutoff, nl["neighbor_list"], nl["neighbor_mask"])
        F = pot.lj_force_nlist(x, BOX10, sigma, eps, cutoff, nl["neighbor_list"], nl["neighbor_mask"])
        assert np.isclose(e, e_ref, rtol=1e-5, atol=1e-8)
        assert np.allclose(F, np.array([f_ref, -f_ref]), rtol=1e-5, atol=1e-5 * max(1.0, np.abs(f_ref).max()))
        assert np.isclose(pot.lj_energy_nopbc(x, sigma, eps, cutoff), e_ref, rtol=1e-5, atol=1e-8)


def _ho_trace(kcal_per_A2, dt, n, refresh):
    k = kcal_per_A2 * 4.184 * 100.0
    x0 = np.zeros((1, 3), f32)
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))
    _, _, _, en = dyn.langevin_run(x0, None, [39.948], 300.0, dt, 1.0, st, n,
                                   lambda x: pot.ho_force(x, x0, k), lambda x: pot.ho_energy(x, x0, k, 0.0),
                                   report_interval=1, refresh_velocities=refresh)
    return np.array(en)


def test_langevin_golden_trace(goldens):
    # chiron/tests/test_mcmc.py:81-84 asserts jnp.allclose (rtol 1e-5, atol 1e-8)
    ref = np.array(goldens["langevin_ho_energy_trace"]["values"])
    assert np.allclose(_ho_trace(100.0, 0.002, 5, True), ref, rtol=1e-5, atol=1e-8)


def test_langevin_golden_trace_20(goldens):
    ref = np.array(goldens["langevin_ho_energy_trace_20"]["values"])
    assert np.allclose(_ho_trace(1.0, 0.001, 20, False), ref, rtol=1e-5, atol=1e-8)


def test_barostat_golden_counts(goldens):
    g = goldens["mc_barostat_counts"]
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], f32)
    box = BOX10.copy()
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))
    red = lambda x, b: dyn.reduced_potential(0.0, 300.0, 1.0, f32(f32(b[0, 0] * b[1, 1]) * b[2, 2]))  # noqa: E731
    u = red(x, box)
    acc = 0
    for _ in range(g["n_proposed"]):
        x, box, u, a = dyn.mc_barostat_step(x, box, st, 0.1, u, red)
        acc += a
    assert acc == g["n_accepted"]
    # beta P V identity of chiron/tests/test_mcmc.py:440-448
    V = float(box[0, 0]) * float(box[1, 1]) * float(box[2, 2])
    assert np.isclose(u, 101325.0 * V * 1e-27 / (dyn.KB_J_PER_K * 300.0), rtol=1e-3)


def test_cell_free_pair_set_matches_bruteforce_distance_definition():
    """The half list is exactly {(i,j): i<j, d_ij < rc+skin}; counts agree with a float64 check away
    from the boundary."""
    rng = np.random.default_rng(0)
    L = 3.0
    x = (rng.random((200, 3)) * L).astype(f32)
    box = np.eye(3, dtype=f32) * L
    out = pairs.build_neighborlist(x, box, 0.8, 0.2, 10)
    keys = pairs.pair_keys(out["neighbor_list"], out["n_neighbors"])
    d = x[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)
    d -= L * np.round(d / L)
    dist = np.sqrt((d ** 2).sum(-1))
    iu = np.triu_indices(200, 1)
    sure_in = dist[iu] < 1.0 - 1e-5
    sure_out = dist[iu] > 1.0 + 1e-5
    k_all = iu[0].astype(np.int64) * 200 + iu[1]
    assert np.all(np.isin(k_all[sure_in], keys))
    assert not np.any(np.isin(k_all[sure_out], keys))
