"""GPU parity tests: every CUDA entry point (called through the C ABI via the Python mirror of
chiron's classes) against the CPU oracle on identical seeded inputs, and against the reference's
golden vectors.  Bar: bit-exact for lists / masks / counts / PRNG bits; energies, forces and
trajectories within rel 1e-5 (the tolerance BASELINE.json states for fp32), written in each test."""
import copy

import numpy as np
import pytest
import torch

from oracle import dynamics as dyn
from oracle import jax_random as jr
from oracle import pairs, potentials as pot

pytestmark = pytest.mark.gpu
f32 = np.float32
BOX10 = np.eye(3, dtype=f32) * 10


def _np(t):
    return t.detach().cpu().numpy()


def _lj_system(n_side, rho_star, seed, sigma=0.34):
    from chiron_b200 import unit
    from chiron_b200.testsystems import LennardJonesFluid
    lj = LennardJonesFluid(nparticles=n_side ** 3, reduced_density=rho_star, sigma=sigma * unit.nanometer, seed=seed)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=f32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=f32)
    return lj, x, box


# ---------------------------------------------------------------------------------------------------
# Space, PRNG
# ---------------------------------------------------------------------------------------------------
def test_space_goldens_and_random_points(cuda_device, goldens):
    from chiron_b200.neighbors import OrthogonalNonPeriodicSpace, OrthogonalPeriodicSpace
    g = goldens["space_periodic"]
    sp, sn = OrthogonalPeriodicSpace(), OrthogonalNonPeriodicSpace()
    r, d = sp.displacement(np.array(g["p1"], f32), np.array(g["p2"], f32), BOX10)
    assert np.array_equal(_np(r), np.array(g["r_ij"], f32)) and np.array_equal(_np(d), np.array(g["dist"], f32))
    for a, b in zip(g["wrap_in"], g["wrap_out"]):
        assert np.array_equal(_np(sp.wrap(np.array(a, f32), BOX10)), np.array(b, f32))
    r, d = sn.displacement(np.array(g["p1"], f32), np.array(g["p2"], f32), BOX10)
    assert np.array_equal(_np(r), np.array([[-1, 0, 0], [-6, 0, 0]], f32))
    assert np.array_equal(_np(sn.wrap(np.array([11, -1, 2], f32), BOX10)), np.array([11, -1, 2], f32))
    with pytest.raises(ValueError):
        sp.displacement(np.zeros((1, 3), f32), np.zeros((1, 3), f32), None)
    # random points, including far outside the box: bit-exact vs the oracle
    rng = np.random.default_rng(3)
    box = np.diag([3.1, 4.7, 2.9]).astype(f32)
    a = (rng.normal(size=(5000, 3)) * 6).astype(f32)
    b = (rng.normal(size=(5000, 3)) * 6).astype(f32)
    r, d = sp.displacement(a, b, box)
    ro, do = pairs.displacement(a, b, box)
    assert np.array_equal(_np(r), ro) and np.array_equal(_np(d), do)
    assert np.array_equal(_np(sp.wrap(a, box)), pairs.wrap(a, box))


def test_threefry_stream_bit_exact(cuda_device):
    from chiron_b200 import random as crandom
    for seed, n in ((0, 1), (1234, 3), (7, 6), (99, 3001), (5, 30000)):
        key = jr.PRNGKey(seed)
        u = crandom.uniform(key, (n,), -1.0, 1.0)
        assert np.array_equal(_np(u), jr.uniform(key, (n,), -1.0, 1.0))
        z = _np(crandom.normal(key, (n,)))
        zo = jr.normal(key, (n,))
        # erfinv goes through log1p whose last bit differs between libm and CUDA: 2 ulp
        assert np.allclose(z, zo, rtol=3e-7, atol=1e-7)
    assert np.isclose(float(crandom.normal(jr.PRNGKey(0), (1,))[0]), -0.20584226, atol=1e-7)
    assert np.isclose(float(crandom.normal(jr.PRNGKey(42), (1,))[0]), -0.18471177, atol=1e-7)


# ---------------------------------------------------------------------------------------------------
# NeighborListNsqrd / PairListNsqrd
# ---------------------------------------------------------------------------------------------------
def _state(x, box, seed=1234):
    from chiron_b200 import unit
    from chiron_b200.states import SamplerState
    from chiron_b200.utils import PRNG
    PRNG.set_seed(seed)
    return SamplerState(positions=unit.Quantity(x, unit.nanometer), current_PRNG_key=PRNG.get_random_key(),
                        box_vectors=None if box is None else unit.Quantity(box, unit.nanometer))


@pytest.mark.parametrize("builder", ["nsq", "cell"])
def test_neighborlist_goldens(cuda_device, goldens, builder):
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    g = goldens["nlist_pair2"]
    x = np.array([[0, 0, 0], [1, 0, 0]], f32)
    state = _state(x, BOX10)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.1 * unit.nanometer, skin=0.1 * unit.nanometer,
                           n_max_neighbors=5, builder=builder)
    nl.build_from_state(state)
    assert nl.is_built and np.array_equal(_np(nl.ref_positions), x) and np.array_equal(_np(nl.box_vectors), BOX10)
    nl.build(state.positions, state.box_vectors)
    assert _np(nl.neighbor_list).tolist() == g["neighbor_list"]
    assert _np(nl.n_neighbors).tolist() == g["n_neighbors"]
    assert _np(nl.neighbor_mask).tolist() == g["neighbor_mask"]
    n, lst, mask, dist, rij = nl.calculate(x)
    assert _np(n).tolist() == [1, 0] and tuple(lst.shape) == (2, 5)
    assert _np(mask).tolist() == g["neighbor_mask"] and np.all(_np(dist) == 1.0)
    assert np.all(_np(rij)[0] == np.array([-1, 0, 0], f32)) and np.all(_np(rij)[1] == np.array([1, 0, 0], f32))
    assert nl.check(torch.as_tensor(x)) is False
    assert nl.check(torch.as_tensor(x + f32(0.1))) is True
    assert nl.check(torch.zeros(3, 3)) is True

    c = goldens["nlist_cube8"]
    grid = np.mgrid[0:2, 0:2, 0:2].astype(f32)
    x8 = np.stack(grid.reshape(3, -1), axis=1).astype(f32)
    state = _state(x8, BOX10)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=2.1 * unit.nanometer, skin=0.1 * unit.nanometer,
                           n_max_neighbors=5, builder=builder)
    nl.build_from_state(state)
    assert _np(nl.n_neighbors).tolist() == c["n_neighbors"]
    assert _np(nl.calculate(x8)[0]).tolist() == c["n_neighbors"]
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.1 * unit.nanometer, skin=1.1 * unit.nanometer,
                           n_max_neighbors=5, builder=builder)
    nl.build_from_state(state)
    n, lst, mask, dist, rij = nl.calculate(x8)
    assert _np(n).tolist() == c["n_interacting_cutoff_1p1"]
    assert tuple(lst.shape) == (8, 17) and _np(lst).tolist() == c["neighbor_list_8x17"]
    # error contract (chiron/tests/test_pairs.py:163-237)
    with pytest.raises(ValueError):
        nl.build_from_state(_state(x8, None))
    with pytest.raises(ValueError):
        nl.build(x8, np.zeros((4, 3), f32))
    with pytest.raises(ValueError):
        nl.build(unit.Quantity(x8, unit.radian), BOX10)
    with pytest.raises(ValueError):
        nl.build(unit.Quantity(x8, unit.nanometer), unit.Quantity(BOX10, unit.radian))


def _build_ref(x, box, cutoff, skin, M):
    return pairs.build_neighborlist(x, box, cutoff, skin, M)


@pytest.mark.parametrize("n_side,rho", [(6, 0.8), (10, 0.8), (10, 0.1)])
def test_neighborlist_bit_exact_vs_oracle(cuda_device, n_side, rho):
    """Lists, masks, counts and n_max growth identical to the oracle; cell builder == N^2 builder."""
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    _, x, box = _lj_system(n_side, rho, seed=11)
    x = x + np.random.default_rng(5).normal(0, 0.05, x.shape).astype(f32)   # some outside the box
    ref = _build_ref(x, box, 1.02, 0.5, 30)
    out = {}
    for builder in ("nsq", "cell"):
        nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=0.5 * unit.nanometer,
                               n_max_neighbors=30, builder=builder)
        nl.build(x, box)
        assert nl.n_max_neighbors == ref["n_max_neighbors"]
        assert np.array_equal(_np(nl.n_neighbors), ref["n_neighbors"])
        assert np.array_equal(_np(nl.neighbor_list).astype(np.uint32), ref["neighbor_list"])
        assert np.array_equal(_np(nl.neighbor_mask), ref["neighbor_mask"])
        out[builder] = nl
    # calculate at displaced positions: bit-exact masks, distances and displacement vectors
    x2 = pairs.wrap(x + np.random.default_rng(6).normal(0, 0.03, x.shape).astype(f32), box)
    n, lst, mask, dist, rij = out["cell"].calculate(x2)
    no, _, mo, do, ro = pairs.calculate_neighborlist(x2, box, 1.02, ref["neighbor_list"], ref["neighbor_mask"])
    assert np.array_equal(_np(n), no) and np.array_equal(_np(mask), mo)
    assert np.array_equal(_np(dist), do) and np.array_equal(_np(rij), ro)
    assert out["cell"].check(torch.as_tensor(x2)) == pairs.check_neighborlist(x2, x, box, 0.5)
    x3 = x.copy()
    x3[7, 0] += f32(0.25)
    assert out["cell"].check(torch.as_tensor(x3)) == pairs.check_neighborlist(x3, x, box, 0.5) is True


def test_neighborlist_nonperiodic_and_empty_rows(cuda_device):
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalNonPeriodicSpace
    rng = np.random.default_rng(2)
    x = (rng.random((300, 3)) * 4).astype(f32)
    x[0] = [50, 50, 50]          # particle 0 isolated: fill value rule (fill == i -> +1)
    ref = pairs.build_neighborlist(x, None, 0.9, 0.3, 12, periodic=False)
    nl = NeighborListNsqrd(OrthogonalNonPeriodicSpace(), cutoff=0.9 * unit.nanometer, skin=0.3 * unit.nanometer,
                           n_max_neighbors=12)
    nl.build(x, None)
    assert np.array_equal(_np(nl.neighbor_list).astype(np.uint32), ref["neighbor_list"])
    assert np.array_equal(_np(nl.n_neighbors), ref["n_neighbors"])
    assert _np(nl.neighbor_list)[0, 0] == 1 and _np(nl.n_neighbors)[0] == 0


def test_pairlist_goldens_and_oracle(cuda_device, goldens):
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace, PairListNsqrd
    g = goldens["pairlist_cube8"]
    grid = np.mgrid[0:2, 0:2, 0:2].astype(f32)
    x8 = np.stack(grid.reshape(3, -1), axis=1).astype(f32)
    pl = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=2.1 * unit.nanometer)
    pl.build_from_state(_state(x8, BOX10))
    assert pl.is_built and _np(pl.all_pairs).tolist() == g["all_pairs"]
    n, ap, mask, dist, rij = pl.calculate(x8)
    assert _np(n).tolist() == [7, 6, 5, 4, 3, 2, 1, 0] and tuple(mask.shape) == (8, 7)
    assert np.allclose(_np(dist), np.array(g["distances"], f32), rtol=0, atol=1e-7)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=2.1 * unit.nanometer, skin=0.1 * unit.nanometer,
                           n_max_neighbors=20)
    nl.build_from_state(_state(x8, BOX10))
    n1, _, m1, d1, _ = nl.calculate(x8)
    assert float(torch.where(mask != 0, dist, 0 * dist).sum()) == float(torch.where(m1 != 0, d1, 0 * d1).sum())
    x2 = np.array([[0, 0, 0], [1, 0, 0]], f32)
    pl = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.1 * unit.nanometer)
    pl.build_from_state(_state(x2, BOX10))
    assert _np(pl.all_pairs).tolist() == [[1], [0]] and _np(pl.reduction_mask).tolist() == [[True], [False]]
    n, ap, mask, dist, rij = pl.calculate(x2)
    assert _np(n).tolist() == [1, 0] and _np(mask).tolist() == [[1], [0]] and _np(dist).tolist() == [[1.0], [1.0]]
    assert _np(rij).tolist() == [[[-1.0, 0.0, 0.0]], [[1.0, 0.0, 0.0]]]
    assert pl.check(torch.as_tensor(x2)) is False and pl.check(torch.zeros(3, 3)) is True
    pl.cutoff = 0.5 * unit.nanometer
    assert _np(pl.calculate(x2)[2]).tolist() == [[0], [0]]
    pl.cutoff = None
    assert _np(pl.calculate(x2)[2]).tolist() == [[1], [0]]
    with pytest.raises(ValueError):
        pl.calculate(np.zeros((3, 3), f32))
    # random system vs oracle, bit exact
    rng = np.random.default_rng(4)
    x = (rng.random((97, 3)) * 3).astype(f32)
    box = np.eye(3, dtype=f32) * 3
    pl = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.0 * unit.nanometer)
    pl.build(x, box)
    apo, redo = pairs.build_pairlist(97)
    no, _, mo, do, ro = pairs.calculate_pairlist(x, box, 1.0, apo, redo)
    n, ap, mask, dist, rij = pl.calculate(x)
    assert np.array_equal(_np(ap).astype(np.uint32), apo) and np.array_equal(_np(pl.reduction_mask), redo)
    assert np.array_equal(_np(n), no) and np.array_equal(_np(mask), mo)
    assert np.array_equal(_np(dist), do) and np.array_equal(_np(rij), ro)


# ---------------------------------------------------------------------------------------------------
# Potentials
# ---------------------------------------------------------------------------------------------------
def test_lj_two_particles_like_reference_test(cuda_device):
    """chiron/tests/test_potential.py:155-230."""
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    sigma, epsilon, cutoff, skin = 1.0, 1.0, 3.0, 0.5
    lj = LJPotential(None, unit.Quantity(sigma, unit.nanometer), unit.Quantity(epsilon, unit.kilojoules_per_mole),
                     unit.Quantity(cutoff, unit.nanometer))
    for i in range(1, 11):
        x = np.array([[0, 0, 0], [i * 0.25 * 2 ** (1 / 6), 0, 0]], f32)
        nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=unit.Quantity(cutoff, unit.nanometer),
                               skin=unit.Quantity(skin, unit.nanometer), n_max_neighbors=5)
        nl.build_from_state(_state(x, BOX10))
        d = float(np.linalg.norm(x[0] - x[1]))
        e_ref = 4.0 * epsilon * ((sigma / d) ** 12 - (sigma / d) ** 6) if d < cutoff else 0.0
        f_ref = (24 * (epsilon / (d * d)) * (2 * (sigma / d) ** 12 - (sigma / d) ** 6)) * (x[0] - x[1]) if d < cutoff else 0 * x[0]
        F_ref = np.array([f_ref, -f_ref])
        assert np.isclose(float(lj.compute_energy(x)), e_ref, rtol=1e-5, atol=1e-8)
        assert np.isclose(float(lj.compute_energy(x, nl)), e_ref, rtol=1e-5, atol=1e-8)
        atol = 1e-5 * max(1.0, float(np.abs(F_ref).max()))
        for F in (lj.compute_force(x), lj.compute_force(x, nl), lj.compute_force_analytical(x)):
            assert np.allclose(_np(F), F_ref, rtol=1e-5, atol=atol)
    with pytest.raises(ValueError):
        nl2 = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=unit.Quantity(cutoff, unit.nanometer),
                                skin=unit.Quantity(skin, unit.nanometer), n_max_neighbors=5)
        lj.compute_energy(x, nl2)                       # not built
    with pytest.raises(ValueError):
        nl3 = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=unit.Quantity(2.0, unit.nanometer),
                                skin=unit.Quantity(skin, unit.nanometer), n_max_neighbors=5)
        nl3.build_from_state(_state(x, BOX10))
        lj.compute_energy(x, nl3)                       # cutoff mismatch


@pytest.mark.parametrize("n_side,rho", [(10, 0.8), (10, 0.1), (16, 0.8)])
def test_lj_fluid_energy_force_vs_oracle(cuda_device, n_side, rho):
    """Energy rel 1e-5; forces within 1e-5 of the force scale (max |F|), both list kinds."""
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace, PairListNsqrd
    from chiron_b200.potential import LJPotential
    lj_sys, x, box = _lj_system(n_side, rho, seed=21)
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    lj = LJPotential(lj_sys.topology, lj_sys.sigma, lj_sys.epsilon, 1.02 * unit.nanometer)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=0.5 * unit.nanometer,
                           n_max_neighbors=180, builder="cell")
    nl.build(x, box)
    x2 = pairs.wrap(x + np.random.default_rng(8).normal(0, 0.02, x.shape).astype(f32), box)
    assert not nl.check(torch.as_tensor(x2))
    e_ref, F_ref, n_int = pot.lj_energy_force_bruteforce(x2, box, sigma, eps, rc)
    scale = float(np.abs(F_ref).max())
    for lst in (nl, ):
        e, F = lj.compute_energy_and_force(x2, lst)
        assert np.isclose(float(e), e_ref, rtol=1e-5)
        assert np.allclose(_np(F), F_ref, rtol=1e-5, atol=1e-5 * scale)
        assert np.isclose(float(lj.compute_energy(x2, lst)), e_ref, rtol=1e-5)
        assert np.allclose(_np(lj.compute_force(x2, lst)), F_ref, rtol=1e-5, atol=1e-5 * scale)
    if n_side <= 10:
        pl = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer)
        pl.build(x, box)
        e, F = lj.compute_energy_and_force(x2, pl)
        assert np.isclose(float(e), e_ref, rtol=1e-5)
        assert np.allclose(_np(F), F_ref, rtol=1e-5, atol=1e-5 * scale)
        # oracle over the reference-shaped list gives the same numbers
        ref = pairs.build_neighborlist(x, box, 1.02, 0.5, 180)
        assert np.isclose(pot.lj_energy_nlist(x2, box, sigma, eps, rc, ref["neighbor_list"], ref["neighbor_mask"]),
                          e_ref, rtol=1e-6)


def test_harmonic_oscillator_goldens(cuda_device, goldens):
    from chiron_b200 import unit
    from chiron_b200.potential import HarmonicOscillatorPotential, IdealGasPotential
    from chiron_b200.testsystems import HarmonicOscillator
    g = goldens["ho_energies"]
    ho = HarmonicOscillator()
    p = HarmonicOscillatorPotential(ho.topology, g["k_kcal_per_mol_A2"] * unit.kilocalories_per_mole / unit.angstroms ** 2,
                                    unit.Quantity(np.array([[0.0, 0.0, 0.0]]), unit.angstrom), 0.0 * unit.kilocalories_per_mole)
    for pos, e in zip(g["positions_A"], g["energies"]):
        x = (np.array([pos]) * 0.1).astype(f32)
        assert np.isclose(float(p.compute_energy(x)), e, rtol=1e-5, atol=1e-8)
    F = p.compute_force(x)
    assert tuple(F.shape) == x.shape and np.allclose(_np(F), pot.ho_force(x, np.zeros((1, 3), f32), p.k), rtol=1e-6)
    ig = IdealGasPotential(ho.topology)
    assert ig.compute_energy(x) == 0.0 and float(ig.compute_force(x).abs().sum()) == 0.0


def test_subset_delta_energy_matches_full_difference(cuda_device):
    from chiron_b200 import _lib
    _, x, box = _lj_system(10, 0.8, seed=31)
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    rng = np.random.default_rng(1)
    for moved in ([17], [3, 500, 999], list(range(40, 60))):
        xn = x.copy()
        xn[moved] += rng.normal(0, 0.01, (len(moved), 3)).astype(f32)
        xn = pairs.wrap(xn, box)
        e0, _, _ = pot.lj_energy_force_bruteforce(x, box, sigma, eps, rc, want_force=False)
        e1, _, _ = pot.lj_energy_force_bruteforce(xn, box, sigma, eps, rc, want_force=False)
        ctx = _lib.get_context()
        xo_t, xn_t = _lib.as_device_f32(x), _lib.as_device_f32(xn)
        ids = torch.as_tensor(moved, dtype=torch.int32, device=xo_t.device)
        delta = torch.zeros((), dtype=torch.float64, device=xo_t.device)
        ctx.call("chx_lj_subset_delta_energy", _lib.ptr(xo_t), _lib.ptr(xn_t), x.shape[0], _lib.ptr(ids), len(moved),
                 float(box[0, 0]), float(box[1, 1]), float(box[2, 2]), 1, sigma, eps, rc, _lib.ptr(delta))
        # fp32 tolerance: rel 1e-5 of the magnitude of the pair terms the moved rows sum over
        # (the difference itself is a small number left after cancellation)
        scale = 0.0
        for xx in (x.astype(np.float64), xn.astype(np.float64)):
            d = xx[moved][:, None, :] - xx[None, :, :]
            d -= np.diag(box) * np.round(d / np.diag(box))
            r = np.sqrt((d * d).sum(-1))
            r = r[(r > 0) & (r < rc)]
            scale += float((4 * eps * ((sigma / r) ** 12 + (sigma / r) ** 6)).sum())
        assert abs(float(delta) - (e1 - e0)) <= 1e-5 * scale


# ---------------------------------------------------------------------------------------------------
# Langevin: golden traces through the GPU path, single step, engine vs building blocks vs oracle
# ---------------------------------------------------------------------------------------------------
def _ho_run(k, dt_fs, nsteps, refresh):
    from chiron_b200 import unit
    from chiron_b200.integrators import LangevinIntegrator
    from chiron_b200.potential import HarmonicOscillatorPotential
    from chiron_b200.reporters import LangevinDynamicsReporter
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import HarmonicOscillator
    from chiron_b200.utils import PRNG
    ho = HarmonicOscillator()
    potential = HarmonicOscillatorPotential(ho.topology, k, U0=ho.U0)
    ts = ThermodynamicState(potential=potential, temperature=300 * unit.kelvin)
    PRNG.set_seed(1234)
    state = SamplerState(positions=ho.positions, current_PRNG_key=PRNG.get_random_key())
    reporter = LangevinDynamicsReporter(f"ho_{nsteps}")
    reporter.reset_reporter_file()
    integ = LangevinIntegrator(timestep=dt_fs * unit.femtosecond, reporter=reporter, report_interval=1,
                               refresh_velocities=refresh)
    integ.run(state, ts, number_of_steps=nsteps)
    return np.asarray(reporter.get_property("potential_energy"), dtype=np.float64).reshape(-1)


def test_langevin_golden_traces_on_gpu(cuda_device, goldens, tmp_path):
    """chiron/tests/test_mcmc.py:12-84 and chiron/tests/test_utils.py:53-113 through the CUDA path."""
    from chiron_b200 import unit
    from chiron_b200.reporters import BaseReporter
    BaseReporter.set_directory(tmp_path)
    e = _ho_run(100.0 * unit.kilocalories_per_mole / unit.angstrom ** 2, 2, 5, True)
    assert np.allclose(e, np.array(goldens["langevin_ho_energy_trace"]["values"]), rtol=1e-5, atol=1e-8)
    e = _ho_run(1.0 * unit.kilocalories_per_mole / unit.angstrom ** 2, 1, 20, False)
    assert np.allclose(e, np.array(goldens["langevin_ho_energy_trace_20"]["values"]), rtol=1e-5, atol=1e-8)


def _lj_langevin_setup(n_side, rho, seed, skin=0.5, builder="cell"):
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG
    lj_sys, x, box = _lj_system(n_side, rho, seed=seed)
    potential = LJPotential(lj_sys.topology, lj_sys.sigma, lj_sys.epsilon, 1.02 * unit.nanometer)
    PRNG.set_seed(1234)
    state = SamplerState(positions=lj_sys.positions, current_PRNG_key=PRNG.get_random_key(),
                         box_vectors=lj_sys.box_vectors)
    ts = ThermodynamicState(potential=potential, temperature=300 * unit.kelvin)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=skin * unit.nanometer,
                           n_max_neighbors=180, builder=builder)
    return lj_sys, x, box, potential, state, ts, nl


def _oracle_langevin(x, box, nsteps, skin=0.5, trace=None):
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    nbr = dyn.OracleNeighborList(box, rc, skin, 180)
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))

    def force(xx):
        return pot.lj_force_nlist(xx, box, sigma, eps, rc, nbr.neighbor_list, nbr.neighbor_mask)

    def energy(xx):
        return pot.lj_energy_nlist(xx, box, sigma, eps, rc, nbr.neighbor_list, nbr.neighbor_mask)
    xo, vo, key, en = dyn.langevin_run(x, None, np.full(x.shape[0], 39.948), 300.0, 0.001, 1.0, st, nsteps,
                                       force, energy, report_interval=1, nbr=nbr, trace=trace)
    return xo, vo, key, np.array(en), nbr, st


@pytest.mark.parametrize("fused", [False, True])
def test_lj_langevin_single_and_few_steps_vs_oracle(cuda_device, fused, tmp_path):
    """With the reference's jax.random stream a BAOAB step agrees within 1e-5 (BASELINE.json); after
    10 steps positions / velocities / reported energies still agree to 1e-5 of their scale."""
    from chiron_b200 import unit
    from chiron_b200.integrators import LangevinIntegrator
    from chiron_b200.reporters import BaseReporter, LangevinDynamicsReporter
    BaseReporter.set_directory(tmp_path)
    for nsteps in (1, 10):
        lj_sys, x, box, potential, state, ts, nl = _lj_langevin_setup(8, 0.8, seed=41)
        reporter = LangevinDynamicsReporter(f"lj_{fused}_{nsteps}")
        reporter.reset_reporter_file()
        integ = LangevinIntegrator(timestep=1.0 * unit.femtosecond, reporter=reporter, report_interval=1)
        integ.use_fused_engine = fused
        out, nl_out = integ.run(state, ts, number_of_steps=nsteps, nbr_list=nl)
        assert integ.last_run_stats["path"] == ("fused" if fused else "blocks")
        xo, vo, key, en, _, st = _oracle_langevin(x, box, nsteps)
        xg, vg = _np(out.positions), _np(out.velocities)
        dx = xg - xo
        dx -= np.diag(box) * np.round(dx / np.diag(box))
        assert np.abs(dx).max() < 1e-5 * float(np.diag(box).max())
        assert np.allclose(vg, vo, rtol=1e-5, atol=1e-5 * float(np.abs(vo).max()))
        assert np.array_equal(np.asarray(out.current_PRNG_key), key)
        assert np.array_equal(np.asarray(out._current_PRNG_key), st.key)
        eg = np.asarray(reporter.get_property("potential_energy"), dtype=np.float64).reshape(-1)
        assert eg.shape == en.shape and np.allclose(eg, en, rtol=1e-5)


def test_fused_engine_anisotropic_box_vs_oracle(cuda_device):
    """The replica shape of config 5 scaled down: a 7 x 7 x 14 lattice (686 particles: the last block is padded)
    in a box with Lx = Ly = Lz / 2, small skin so the tables and the reference list are rebuilt on the way; 40 fused
    steps against the oracle."""
    from chiron_b200 import unit
    from chiron_b200.integrators import LangevinIntegrator
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import LennardJonesFluid
    from chiron_b200.utils import PRNG
    nsteps, skin = 40, 0.03
    lj_sys = LennardJonesFluid(cells=(7, 7, 14), reduced_density=0.8, sigma=0.34 * unit.nanometer, seed=47)
    x = np.asarray(lj_sys.positions.value_in_unit(unit.nanometer), dtype=f32)
    box = np.asarray(lj_sys.box_vectors.value_in_unit(unit.nanometer), dtype=f32)
    L = np.diag(box)
    assert x.shape[0] == 686 and abs(L[2] / L[0] - 2.0) < 1e-6 and L[0] > 2 * (1.02 + skin)
    potential = LJPotential(lj_sys.topology, lj_sys.sigma, lj_sys.epsilon, 1.02 * unit.nanometer)
    PRNG.set_seed(1234)
    state = SamplerState(positions=lj_sys.positions, current_PRNG_key=PRNG.get_random_key(),
                         box_vectors=lj_sys.box_vectors)
    ts = ThermodynamicState(potential=potential, temperature=300 * unit.kelvin)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=skin * unit.nanometer,
                           n_max_neighbors=180)
    integ = LangevinIntegrator(timestep=2.0 * unit.femtosecond)
    out, nl_out = integ.run(state, ts, number_of_steps=nsteps, nbr_list=nl)
    assert integ.last_run_stats["path"] == "fused"
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    nbr = dyn.OracleNeighborList(box, rc, skin, 180)
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))
    force = lambda xx: pot.lj_force_nlist(xx, box, sigma, eps, rc, nbr.neighbor_list, nbr.neighbor_mask)  # noqa: E731
    xo, vo, key, _ = dyn.langevin_run(x, None, np.full(x.shape[0], 39.948), 300.0, 0.002, 1.0, st, nsteps, force, nbr=nbr)
    assert nbr.n_builds >= 2 and nl_out.n_builds == nbr.n_builds
    dx = _np(out.positions) - xo
    dx -= L * np.round(dx / L)
    assert np.abs(dx).max() < 5e-5 * float(L.max())
    assert np.allclose(_np(out.velocities), vo, rtol=0, atol=5e-4 * float(np.abs(vo).max()))
    assert np.array_equal(np.asarray(out.current_PRNG_key), key)


def test_fused_engine_rebuild_events_match_oracle(cuda_device):
    """Small skin so several rebuilds happen: the engine's reference-rebuild bookkeeping (count and
    final ref_positions), the lazily materialised list, and the trajectory all follow the oracle."""
    from chiron_b200 import unit
    from chiron_b200.integrators import LangevinIntegrator
    nsteps, skin = 60, 0.02
    lj_sys, x, box, potential, state, ts, nl = _lj_langevin_setup(8, 0.8, seed=43, skin=skin)
    integ = LangevinIntegrator(timestep=2.0 * unit.femtosecond)
    out, nl_out = integ.run(state, ts, number_of_steps=nsteps, nbr_list=nl)
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    nbr = dyn.OracleNeighborList(box, rc, skin, 180)
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))
    force = lambda xx: pot.lj_force_nlist(xx, box, sigma, eps, rc, nbr.neighbor_list, nbr.neighbor_mask)  # noqa: E731
    xo, vo, key, _ = dyn.langevin_run(x, None, np.full(x.shape[0], 39.948), 300.0, 0.002, 1.0, st, nsteps, force, nbr=nbr)
    assert nbr.n_builds >= 3
    assert integ.last_run_stats["reference_rebuilds"] == nbr.n_builds - 1
    assert nl_out.n_builds == nbr.n_builds
    dref = _np(nl_out.ref_positions) - nbr.ref
    assert np.abs(dref).max() < 2e-4        # same rebuild step, positions equal to trajectory tolerance
    dx = _np(out.positions) - xo
    dx -= np.diag(box) * np.round(dx / np.diag(box))
    assert np.abs(dx).max() < 5e-5 * float(np.diag(box).max())
    # the list handed back is the list of its reference positions
    ref = pairs.build_neighborlist(_np(nl_out.ref_positions), box, rc, skin, 180)
    assert np.array_equal(_np(nl_out.n_neighbors), ref["n_neighbors"])
    assert np.array_equal(_np(nl_out.neighbor_list).astype(np.uint32), ref["neighbor_list"])


def test_fused_engine_force_and_energy_vs_bruteforce(cuda_device):
    """Engine tables (Morton sort + tiles + bitmasks): forces/energy of the uploaded state equal the
    all-pairs oracle; candidate-pair count equals the reference list size."""
    from chiron_b200._engine import LJLangevinEngine
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    for n_side, rho in ((8, 0.8), (12, 0.8), (10, 0.1), (16, 0.8)):
        _, x, box = _lj_system(n_side, rho, seed=51)
        n = x.shape[0]
        eng = LJLangevinEngine(n, np.diag(box), sigma, eps, rc, 0.5, 0.001, 1.0, 2.494)
        eng.set_state(x, np.zeros_like(x), np.full(n, 39.948, f32))
        xs, vs, F, ref = eng.get_state(want_force=True, want_ref=True)
        assert np.array_equal(_np(xs), x) and np.array_equal(_np(ref), x)
        e_ref, F_ref, n_int = pot.lj_energy_force_bruteforce(x, box, sigma, eps, rc)
        e = float(eng.energy()[0])
        st = eng.stats()
        assert np.isclose(e, e_ref, rtol=1e-5)
        assert np.allclose(_np(F), F_ref, rtol=1e-5, atol=1e-5 * float(np.abs(F_ref).max()))
        assert st["interacting_pairs"] == n_int
        # the engine's own tables are conservative (fast predicate with a safety margin) on the
        # internal skin: at least every pair within rc + skin_int, nothing beyond the margin
        for skin_int in (0.5, 0.12):
            e2 = LJLangevinEngine(n, np.diag(box), sigma, eps, rc, 0.5, 0.001, 1.0, 2.494, internal_skin=skin_int)
            e2.set_state(x, np.zeros_like(x), np.full(n, 39.948, f32))
            cand = e2.stats()["candidate_pairs"]
            lo = sum(r.size for r in pairs.neighbor_rows(x, box, rc + skin_int))
            hi = sum(r.size for r in pairs.neighbor_rows(x, box, (rc + skin_int) * (1 + 4e-5) + 2e-6))
            assert lo <= cand <= hi
            _, _, F2, _ = e2.get_state(want_force=True)
            assert np.allclose(_np(F2), F_ref, rtol=1e-5, atol=1e-5 * float(np.abs(F_ref).max()))
            e2.close()
        eng.close()


def test_fused_engine_batched_replicas(cuda_device):
    """R replicas in one launch == R independent single-replica runs."""
    from chiron_b200._engine import LJLangevinEngine
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    R = 3
    xs, keys, kts = [], [], [2.0, 2.494, 3.1]
    for r in range(R):
        _, x, box = _lj_system(8, 0.8, seed=60 + r)
        xs.append(x)
        keys.append(jr.PRNGKey(100 + r))
    n = xs[0].shape[0]
    mass = np.full(n, 39.948, f32)
    v0 = [dyn.maxwell_boltzmann(jr.PRNGKey(7 + r), mass, 300.0) for r in range(R)]
    eng = LJLangevinEngine(n, np.diag(box), sigma, eps, rc, 0.05, 0.002, 1.0, 2.494, n_replicas=R)
    eng.set_state(np.stack(xs), np.stack(v0), mass, kts)
    kout, en = eng.run(40, np.stack(keys), report_interval=10)
    xb, vb, _, _ = eng.get_state()
    eng.close()
    for r in range(R):
        e1 = LJLangevinEngine(n, np.diag(box), sigma, eps, rc, 0.05, 0.002, 1.0, kts[r])
        e1.set_state(xs[r], v0[r], mass, [kts[r]])
        k1, en1 = e1.run(40, keys[r], report_interval=10)
        x1, v1, _, _ = e1.get_state()
        # rebuild points are shared by the replicas of one launch, so the fp32 summation order (not
        # the pair set) can differ from a single-replica run: equal to rounding, keys bit-exact
        assert np.allclose(_np(xb)[r], _np(x1), rtol=0, atol=2e-6) and np.allclose(_np(vb)[r], _np(v1), rtol=0, atol=2e-5)
        assert np.array_equal(kout[r], k1[0]) and np.allclose(_np(en)[:, r], _np(en1)[:, 0], rtol=1e-6)
        e1.close()


@pytest.mark.parametrize("mode", ["proactive", "lookahead", "persistent", "shifted_grid"])
def test_fused_engine_batched_replicas_step_loops(cuda_device, monkeypatch, mode):
    """The three step loops a batch of replicas can take without energy reports -- fresh tables for everybody at
    every 50-step chunk (default), halt-driven rebuilds per replica behind the look-ahead loop (CHX_MD_PROACTIVE=0)
    and the persistent kernel -- against independent single-replica runs: keys bit-equal, trajectories to rounding.
    Small internal skin and a hot replica, so that replicas go stale at different steps inside a chunk."""
    from chiron_b200._engine import LJLangevinEngine
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    R, nsteps = 3, 61
    xs, keys, kts = [], [], [1.5, 2.494, 6.0]
    for r in range(R):
        _, x, box = _lj_system(8, 0.8, seed=60 + r)
        xs.append(x)
        keys.append(jr.PRNGKey(100 + r))
    n = xs[0].shape[0]
    mass = np.full(n, 39.948, f32)
    v0 = [dyn.maxwell_boltzmann(jr.PRNGKey(7 + r), mass, 300.0 * kts[r] / 2.494) for r in range(R)]
    if mode == "lookahead":
        monkeypatch.setenv("CHX_MD_PROACTIVE", "0")
    if mode == "persistent":
        monkeypatch.setenv("CHX_MD_PERSIST", "1")
    eng = LJLangevinEngine(n, np.diag(box), sigma, eps, rc, 0.3, 0.002, 1.0, 2.494, n_replicas=R, internal_skin=0.04)
    if mode == "shifted_grid":
        # chx_ljmd_set_chunk_phase: the first chunk after set_state is 50 * (1 - 3/5) = 20 steps long, later
        # chunks 50; the chunk grid carries over from run to run
        eng.set_chunk_phase(3, 5)
    eng.set_state(np.stack(xs), np.stack(v0), mass, kts)
    kout, _ = eng.run(nsteps, np.stack(keys))
    xb, vb, _, _ = eng.get_state()
    assert eng.stats()["table_rebuilds"] >= 3
    if mode == "shifted_grid":
        with pytest.raises(Exception):
            eng.set_chunk_phase(5, 5)
    eng.close()
    monkeypatch.delenv("CHX_MD_PROACTIVE", raising=False)
    monkeypatch.delenv("CHX_MD_PERSIST", raising=False)
    L = np.diag(box)
    for r in range(R):
        e1 = LJLangevinEngine(n, np.diag(box), sigma, eps, rc, 0.3, 0.002, 1.0, kts[r], internal_skin=0.04)
        e1.set_state(xs[r], v0[r], mass, [kts[r]])
        k1, _ = e1.run(nsteps, keys[r])
        x1, v1, _, _ = e1.get_state()
        dx = _np(xb)[r] - _np(x1)
        dx -= L * np.round(dx / L)
        assert np.abs(dx).max() < 2e-5 and np.allclose(_np(vb)[r], _np(v1), rtol=0, atol=5e-4)
        assert np.array_equal(kout[r], k1[0])
        e1.close()


# ---------------------------------------------------------------------------------------------------
# Monte Carlo moves
# ---------------------------------------------------------------------------------------------------
def test_mc_barostat_golden(cuda_device, goldens, tmp_path):
    """chiron/tests/test_mcmc.py:340-452: PE == 0, beta P V identity, n_accepted == 8 of 10."""
    from chiron_b200 import unit
    from chiron_b200.mcmc import MonteCarloBarostatMove
    from chiron_b200.neighbors import OrthogonalPeriodicSpace, PairListNsqrd
    from chiron_b200.potential import IdealGasPotential
    from chiron_b200.reporters import BaseReporter, MCReporter
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import _topology
    from chiron_b200.utils import PRNG
    BaseReporter.set_directory(tmp_path)
    reporter = MCReporter(1)
    move = MonteCarloBarostatMove(volume_max_scale=0.1, number_of_moves=10, reporter=reporter, report_interval=1)
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], f32)
    PRNG.set_seed(1234)
    state = SamplerState(positions=x * unit.nanometer, box_vectors=BOX10 * unit.nanometer,
                         current_PRNG_key=PRNG.get_random_key())
    ts = ThermodynamicState(potential=IdealGasPotential(_topology(8)), temperature=300 * unit.kelvin,
                            pressure=1.0 * unit.atmosphere)
    nl = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=0 * unit.nanometer)
    state, ts, nl = move.update(state, ts, nl)
    pe = reporter.get_property("potential_energy")
    vol = reporter.get_property("volume")
    assert pe[0] == 0 and pe[-1] == 0
    assert np.isclose(float(ts.get_reduced_potential(state)),
                      float(ts.pressure * ts.beta * (float(vol[-1]) * unit.nanometer ** 3)), rtol=1e-3)
    g = goldens["mc_barostat_counts"]
    assert move.statistics["n_proposed"] == g["n_proposed"]
    assert move.statistics["n_accepted"] == g["n_accepted"]


def test_mc_displacement_lj_vs_oracle(cuda_device):
    """MonteCarloDisplacementMove on an LJ fluid: the accept/reject sequence, final positions and key
    follow the oracle (decisions are identical unless a uniform lands within fp32 noise of the
    threshold, which this seed does not hit)."""
    from chiron_b200 import unit
    from chiron_b200.mcmc import MonteCarloDisplacementMove
    lj_sys, x, box, potential, state, ts, nl = _lj_langevin_setup(6, 0.8, seed=71, builder="nsq")
    nl.build_from_state(state)
    move = MonteCarloDisplacementMove(displacement_sigma=0.0005 * unit.nanometer, number_of_moves=12)
    out, _, nl_out = move.update(state, ts, nl)
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    nbr = dyn.OracleNeighborList(box, rc, 0.5, 180)
    nbr.build(x)
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))
    red = lambda xx, b: dyn.reduced_potential(  # noqa: E731
        pot.lj_energy_nlist(xx, box, sigma, eps, rc, nbr.neighbor_list, nbr.neighbor_mask), 300.0)
    xo, u, acc = x, red(x, box), 0
    for _ in range(12):
        xo, u, a = dyn.mc_displacement_step(xo, box, st, 0.0005, u, red, nbr=nbr)
        acc += a
    assert 0 < acc < 12
    assert move.statistics == dict(n_accepted=acc, n_proposed=12)
    assert np.allclose(_np(out.positions), xo, rtol=0, atol=2e-6)
    assert np.array_equal(np.asarray(out._current_PRNG_key), st.key)


def test_mc_displacement_subset_delta_path(cuda_device):
    """atom_subset moves through the single-pass delta-energy kernel == full re-evaluation path."""
    from chiron_b200 import unit
    from chiron_b200.mcmc import MonteCarloDisplacementMove
    res = []
    for use_delta in (True, False):
        lj_sys, x, box, potential, state, ts, nl = _lj_langevin_setup(8, 0.8, seed=73, builder="cell")
        nl.build_from_state(state)
        move = MonteCarloDisplacementMove(displacement_sigma=0.01 * unit.nanometer, number_of_moves=30,
                                          atom_subset=[5], use_delta_energy=use_delta)
        out, _, _ = move.update(state, ts, nl)
        res.append((move.statistics["n_accepted"], _np(out.positions)))
    assert res[0][0] == res[1][0] and 0 < res[0][0] < 30
    assert np.array_equal(res[0][1], res[1][1])
    moved = np.nonzero(np.any(res[0][1] != x, axis=1))[0]
    assert moved.tolist() == [5]


def test_persistent_step_kernel_matches_per_launch_path_and_odd_run_lengths(cuda_device, monkeypatch):
    """Engine v5 runs the step loop in ONE persistent cooperative kernel (k_md_steps: pieces of the tile
    stream per warp, split blocks finished by the last part to arrive, grid barrier between steps); the
    per-launch path (CHX_MD_PERSIST=0: k_md_force<UPDATE>, CUDA-graph replays) evaluates the same pairs with
    the partial forces of a block added in a different order.  Odd run lengths (75, 40: the state ends in
    buffer B and is copied back), table rebuilds at odd steps: keys and rebuild bookkeeping identical,
    trajectories equal to rounding; and persistent 75 + 40 == persistent 115 bit for bit."""
    from chiron_b200 import random as crandom
    from chiron_b200._engine import LJLangevinEngine
    lj_sys, x, box = _lj_system(12, 0.8, seed=91)
    n = x.shape[0]
    rng = np.random.default_rng(5)
    v = rng.normal(0, 0.25, (n, 3)).astype(f32)
    mass = np.full(n, 39.948, f32)
    out = {}
    for mode, runs in (("persist", (75, 40)), ("launch", (75, 40)), ("persist_single", (115,))):
        monkeypatch.setenv("CHX_MD_PERSIST", "0" if mode == "launch" else "1")
        eng = LJLangevinEngine(n, np.diag(box), 0.34, 0.238 * 4.184, 1.02, 0.3, 0.002, 1.0, 2.494,
                               internal_skin=0.05, device=cuda_device)
        eng.set_state(x, v, mass, [2.494])
        keys = crandom.PRNGKey(5).reshape(1, 2)
        for k in runs:
            keys, _ = eng.run(k, keys)
        xs, vs, _, ref = eng.get_state(want_ref=True)
        st = eng.stats()
        out[mode] = (_np(xs), _np(vs), np.array(keys), st["table_rebuilds"], st["reference_rebuilds"], _np(ref))
        eng.close()
    a, b, c = out["persist"], out["launch"], out["persist_single"]
    assert a[3] >= 3 and abs(a[3] - b[3]) <= 1 and a[4] == b[4]
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[2], c[2])
    L = np.diag(box)
    dx = a[0] - b[0]
    dx -= L * np.round(dx / L)
    assert np.abs(dx).max() < 1e-4 and np.allclose(a[1], b[1], rtol=0, atol=2e-3)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1]) and np.array_equal(a[5], c[5])
    assert a[3] == c[3] and a[4] == c[4]


# ---------------------------------------------------------------------------------------------------
# device-resident Metropolis loop (csrc/mc.cu) == the reference-shaped step-by-step path
# ---------------------------------------------------------------------------------------------------
def _run_move_both_ways(make, n_moves, **move_kw):
    """Run the same MonteCarloDisplacementMove through the device loop and step by step."""
    from chiron_b200.mcmc import MonteCarloDisplacementMove
    out = {}
    for loop in (True, False):
        state, ts, nl = make()
        move = MonteCarloDisplacementMove(number_of_moves=n_moves, **copy.deepcopy(move_kw))  # autotune mutates sigma
        move.device_loop = loop
        s, _, nl_out = move.update(state, ts, nl)
        out[loop] = (move.statistics, _np(s.positions), np.asarray(s._current_PRNG_key).copy(),
                     move.number_of_attemps_made, move.displacement_sigma, nl_out)
    return out


@pytest.mark.parametrize("subset", [None, [5]])
def test_mc_device_loop_matches_stepwise_lj(cuda_device, subset):
    """LJ + NeighborListNsqrd: identical decisions, positions and key (the two paths share the proposal
    arithmetic bit for bit; the reduced potentials differ by fp32 rounding only)."""
    from chiron_b200 import unit

    def make():
        lj_sys, x, box, potential, state, ts, nl = _lj_langevin_setup(6, 0.8, seed=71, builder="nsq")
        nl.build_from_state(state)
        return state, ts, nl
    sig = 0.0005 if subset is None else 0.02
    out = _run_move_both_ways(make, 60, displacement_sigma=sig * unit.nanometer, atom_subset=subset)
    a, b = out[True], out[False]
    assert a[0] == b[0] and 0 < a[0]["n_accepted"] < 60 and a[0]["n_proposed"] == 60
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2])
    assert a[3] == b[3] == 60


def test_mc_device_loop_halts_for_list_rebuilds_and_autotunes(cuda_device):
    """A skin small enough that proposals trigger NeighborListNsqrd.check(): the loop stops before
    those moves, they run through the building-block path (rebuild), and the sequence still equals the
    step-by-step one; autotune fires at the same attempts."""
    from chiron_b200 import unit

    def make():
        lj_sys, x, box, potential, state, ts, nl = _lj_langevin_setup(6, 0.8, seed=72, skin=0.004, builder="cell")
        nl.build_from_state(state)
        return state, ts, nl
    out = _run_move_both_ways(make, 45, displacement_sigma=0.0005 * unit.nanometer, autotune=True,
                              autotune_interval=10)
    a, b = out[True], out[False]
    assert a[0] == b[0] and 0 < a[0]["n_accepted"] < 45
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert float(a[4].value_in_unit(unit.nanometer)) == float(b[4].value_in_unit(unit.nanometer))
    assert a[5].n_builds == b[5].n_builds and a[5].n_builds > 1
    assert np.array_equal(_np(a[5].ref_positions), _np(b[5].ref_positions))


def test_mc_device_loop_ho_and_ideal_gas(cuda_device):
    """Harmonic oscillator (no list) and ideal gas (pair list, NPT reduced potential) displacement moves."""
    from chiron_b200 import unit
    from chiron_b200.neighbors import OrthogonalPeriodicSpace, PairListNsqrd
    from chiron_b200.potential import HarmonicOscillatorPotential, IdealGasPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import _topology
    from chiron_b200.utils import PRNG

    def make_ho():
        PRNG.set_seed(1234)
        pot_ho = HarmonicOscillatorPotential(_topology(5), k=100.0 * unit.kilocalories_per_mole / unit.angstrom ** 2,
                                             x0=np.zeros((5, 3)) * unit.angstrom)
        x = np.linspace(-0.01, 0.01, 15).reshape(5, 3).astype(f32)
        state = SamplerState(positions=x * unit.nanometer, current_PRNG_key=PRNG.get_random_key())
        return state, ThermodynamicState(pot_ho, temperature=300 * unit.kelvin), None
    out = _run_move_both_ways(make_ho, 80, displacement_sigma=0.01 * unit.angstrom)
    a, b = out[True], out[False]
    assert a[0] == b[0] and 0 < a[0]["n_accepted"] < 80
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])

    def make_ig():
        PRNG.set_seed(1234)
        rng = np.random.default_rng(3)
        x = (rng.random((216, 3)) * 10).astype(f32)
        state = SamplerState(positions=x * unit.nanometer, box_vectors=BOX10 * unit.nanometer,
                             current_PRNG_key=PRNG.get_random_key())
        ts = ThermodynamicState(IdealGasPotential(_topology(216)), temperature=298 * unit.kelvin,
                                pressure=1.0 * unit.atmosphere)
        nl = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=0 * unit.nanometer)
        nl.build_from_state(state)
        return state, ts, nl
    out = _run_move_both_ways(make_ig, 25, displacement_sigma=0.1 * unit.nanometer)
    a, b = out[True], out[False]
    assert a[0] == b[0] == dict(n_accepted=25, n_proposed=25)     # U = 0: every proposal is accepted
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[1].min() >= 0 and a[1].max() < 10                      # wrapped into the box


# ---------------------------------------------------------------------------------------------------
# BASELINE.json's full size (config 4, N = 262,144): size-independent properties
# ---------------------------------------------------------------------------------------------------
def test_full_size_engine_properties(cuda_device):
    """At N = 262,144 the all-pairs oracle is out of reach, so the fused engine is checked through
    properties: (1) its forces / energy / interacting-pair count equal the independent reference-shaped
    path (cell-list NeighborListNsqrd arrays + exact-predicate k_lj_nlist) within rel 1e-5, (2) Newton's
    third law: the forces sum to zero, (3) a run of 64 + 37 steps equals one run of 101 steps bit for
    bit (graph replays, odd run lengths, table rebuilds), particle ids are a permutation, positions stay
    inside the box, and the kinetic temperature stays physical."""
    from chiron_b200 import random as crandom, unit
    from chiron_b200._engine import LJLangevinEngine
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.utils import initialize_velocities
    sigma, eps, rc, skin = 0.34, 0.238 * 4.184, 1.02, 0.5
    lj_sys, x, box = _lj_system(64, 0.8, seed=4)
    n = x.shape[0]
    assert n == 262144
    L = np.diag(box)
    mass = np.full(n, 39.948, f32)
    v0 = _np(initialize_velocities(300 * unit.kelvin, lj_sys.topology, crandom.PRNGKey(11)).value_in_unit_system(
        unit.md_unit_system))
    eng = LJLangevinEngine(n, L, sigma, eps, rc, skin, 0.001, 1.0, 2.494, device=cuda_device)
    eng.set_state(x, v0, mass, [2.494])
    _, _, F, _ = eng.get_state(want_force=True)
    e = float(eng.energy()[0])
    p_int = eng.stats()["interacting_pairs"]
    # (1) independent path
    potential = LJPotential(lj_sys.topology, lj_sys.sigma, lj_sys.epsilon, rc * unit.nanometer)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer,
                           n_max_neighbors=400, builder="cell")
    nl.build(x, box)
    e_ref, F_ref = potential.compute_energy_and_force(x, nl)
    n_nb, _, mask, _, _ = nl.calculate(torch.as_tensor(x, device=cuda_device))
    assert int(mask.sum().item()) == p_int
    assert np.isclose(e, float(e_ref), rtol=1e-5)
    Fn, Frn = _np(F), _np(F_ref)
    assert np.allclose(Fn, Frn, rtol=1e-5, atol=1e-5 * float(np.abs(Frn).max()))
    del nl, mask
    # (2) Newton's third law (fp32 sums of ~90 terms per particle, fp64 total)
    assert np.abs(Fn.astype(np.float64).sum(axis=0)).max() < 1e-6 * np.abs(Fn).sum()
    # (3) split run == single run
    key = crandom.PRNGKey(1234).reshape(1, 2)
    k1, _ = eng.run(64, key)
    k1, _ = eng.run(37, k1)
    xa, va, _, _ = eng.get_state()
    rebuilds = eng.stats()["table_rebuilds"]
    eng.close()
    eng = LJLangevinEngine(n, L, sigma, eps, rc, skin, 0.001, 1.0, 2.494, device=cuda_device)
    eng.set_state(x, v0, mass, [2.494])
    k2, _ = eng.run(101, key)
    xb, vb, _, _ = eng.get_state()
    eng.close()
    assert rebuilds >= 2
    assert np.array_equal(k1, k2)
    assert torch.equal(xa, xb) and torch.equal(va, vb)
    xan = _np(xa)
    assert np.isfinite(xan).all() and (xan >= 0).all() and (xan < L).all()
    kin = 0.5 * 39.948 * (_np(va).astype(np.float64) ** 2).sum()
    T = 2.0 * kin / (3 * n) / 8.314462618e-3
    assert 150.0 < T < 450.0, T     # melting lattice: kinetic energy flows into the potential energy at first


@pytest.mark.parametrize("with_pairlist", [False, True])
def test_mc_device_loop_lj_all_pairs(cuda_device, with_pairlist):
    """LJ without a neighbour list (non-periodic N^2 pairs, potential.py:235-258) and over a periodic
    PairListNsqrd: the device loop's all-pairs energy kernel against the step-by-step path."""
    from chiron_b200 import unit
    from chiron_b200.neighbors import OrthogonalPeriodicSpace, PairListNsqrd
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG

    def make():
        lj_sys, x, box = _lj_system(5, 0.6, seed=81)
        potential = LJPotential(lj_sys.topology, lj_sys.sigma, lj_sys.epsilon, 1.02 * unit.nanometer)
        PRNG.set_seed(1234)
        state = SamplerState(positions=lj_sys.positions, current_PRNG_key=PRNG.get_random_key(),
                             box_vectors=lj_sys.box_vectors)
        ts = ThermodynamicState(potential=potential, temperature=300 * unit.kelvin)
        nl = None
        if with_pairlist:
            nl = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer)
            nl.build_from_state(state)
        return state, ts, nl
    out = _run_move_both_ways(make, 40, displacement_sigma=0.003 * unit.nanometer)
    a, b = out[True], out[False]
    assert a[0] == b[0] and 0 < a[0]["n_accepted"] < 40
    assert np.allclose(a[1], b[1], rtol=0, atol=1e-6) and np.array_equal(a[2], b[2])


# ---------------------------------------------------------------------------------------------------
# barostat loop on the device (csrc/mc.cu, chx_mc_barostat_run) == the step-by-step path
# ---------------------------------------------------------------------------------------------------
def _run_barostat_both_ways(skin, n_moves, scale, autotune=False):
    from chiron_b200 import unit
    from chiron_b200.mcmc import MonteCarloBarostatMove
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG
    out = {}
    for loop in (True, False):
        lj_sys, x, box = _lj_system(12, 0.8, seed=91)
        potential = LJPotential(lj_sys.topology, lj_sys.sigma, lj_sys.epsilon, 1.02 * unit.nanometer)
        PRNG.set_seed(1234)
        state = SamplerState(positions=lj_sys.positions, current_PRNG_key=PRNG.get_random_key(),
                             box_vectors=lj_sys.box_vectors)
        ts = ThermodynamicState(potential=potential, temperature=300 * unit.kelvin, pressure=2000.0 * unit.atmosphere)
        nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=skin * unit.nanometer,
                               n_max_neighbors=180, builder="cell")
        nl.build_from_state(state)
        move = MonteCarloBarostatMove(volume_max_scale=scale, number_of_moves=n_moves, autotune=autotune,
                                      autotune_interval=10)
        move.device_loop = loop
        s, ts_out, nl_out = move.update(state, ts, nl)
        out[loop] = dict(stats=move.statistics, x=_np(s.positions), key=np.asarray(s._current_PRNG_key).copy(),
                         box=_np(s.box_vectors), scale=move.volume_max_scale, nl=nl_out,
                         volume=float(ts_out.volume.value_in_unit(unit.nanometer ** 3)),
                         u=float(ts_out.get_reduced_potential(s, nl_out)), x0=x, box0=box)
    return out


def test_mc_barostat_device_loop_matches_stepwise(cuda_device):
    """Same decisions, key and autotuned scale; positions and box equal to fp32 rounding (the loop takes
    the cube root with CUDA's powf, the host path with NumPy's); the list handed back is the reference
    list of the final configuration."""
    out = _run_barostat_both_ways(skin=0.3, n_moves=30, scale=0.002, autotune=True)
    a, b = out[True], out[False]
    assert a["stats"] == b["stats"] and 0 < a["stats"]["n_accepted"] < 30
    assert np.array_equal(a["key"], b["key"]) and a["scale"] == b["scale"]
    assert np.allclose(a["box"], b["box"], rtol=1e-6) and not np.allclose(a["box"], a["box0"], rtol=1e-5)
    assert np.allclose(a["x"], b["x"], rtol=2e-6, atol=1e-6)
    assert np.isclose(a["volume"], b["volume"], rtol=1e-6) and np.isclose(a["u"], b["u"], rtol=1e-5)
    # the list of the loop == a fresh reference build on the loop's final state
    L = np.diag(a["box"])
    ref = pairs.build_neighborlist(a["x"], a["box"], 1.02, 0.3, int(a["nl"].n_max_neighbors))
    assert np.array_equal(_np(a["nl"].n_neighbors), ref["n_neighbors"])
    assert np.array_equal(_np(a["nl"].neighbor_list).astype(np.uint32), ref["neighbor_list"])
    assert np.array_equal(_np(a["nl"].ref_positions), a["x"]) and L.min() > 3.9


def test_mc_barostat_device_loop_halts_when_the_cell_grid_breaks(cuda_device):
    """List radius just under a third of the box: every compression leaves fewer than 3 cells per edge,
    the loop stops before those moves and they run through the building blocks (O(N^2) builder)."""
    lj_sys, x, box = _lj_system(12, 0.8, seed=91)
    skin = float(box[0, 0]) / 3.0005 - 1.02
    out = _run_barostat_both_ways(skin=skin, n_moves=16, scale=0.002)
    a, b = out[True], out[False]
    assert a["stats"] == b["stats"] and 0 < a["stats"]["n_accepted"] < 16
    assert np.array_equal(a["key"], b["key"])
    assert np.allclose(a["box"], b["box"], rtol=1e-6) and np.allclose(a["x"], b["x"], rtol=2e-6, atol=1e-6)


def test_mc_device_loop_nan_energy_is_rejected_with_one_key_split(cuda_device):
    """Two particles on top of each other: every proposal has a NaN energy, which the reference rejects
    WITHOUT drawing the acceptance uniform (mcmc.py:417-430), so the key advances by one split per move."""
    from chiron_b200 import unit
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import _topology
    from chiron_b200.utils import PRNG

    def make():
        x = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.5, 0.1, 0.0], [0.0, 0.6, 0.2]], f32)
        potential = LJPotential(_topology(4), 0.34 * unit.nanometer, 0.238 * unit.kilocalories_per_mole,
                                1.02 * unit.nanometer)
        PRNG.set_seed(1234)
        state = SamplerState(positions=x * unit.nanometer, current_PRNG_key=PRNG.get_random_key())
        return state, ThermodynamicState(potential=potential, temperature=300 * unit.kelvin), None
    out = _run_move_both_ways(make, 12, displacement_sigma=0.01 * unit.nanometer, atom_subset=[2])
    a, b = out[True], out[False]
    assert a[0] == b[0] == dict(n_accepted=0, n_proposed=12)
    assert np.array_equal(a[2], b[2])
    key = jr.PRNGKey(1234)
    key = jr.split(key)[1]                       # PRNG.get_random_key() hands out the subkey
    for _ in range(12):
        key = jr.split(key)[0]
    assert np.array_equal(a[2], key)


def test_mc_loops_at_config3_size(cuda_device):
    """BASELINE.json config 3 (LJ NPT, N = 32,768, UA-TraPPE methane parameters of Examples/LJ_MCMC.py):
    a few displacement and barostat moves through the device loops against the step-by-step path."""
    from chiron_b200 import unit
    from chiron_b200.mcmc import MonteCarloBarostatMove, MonteCarloDisplacementMove
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import LennardJonesFluid
    from chiron_b200.utils import PRNG
    sigma, eps, rc, skin = 0.373, 0.2941, 1.4, 0.5
    res = {}
    for loop in (True, False):
        lj = LennardJonesFluid(nparticles=32 ** 3, reduced_density=14.08 * sigma ** 3, sigma=sigma * unit.nanometer,
                               epsilon=eps * unit.kilocalories_per_mole, mass=16.04, seed=3, symbol="C")
        pot_ = LJPotential(lj.topology, lj.sigma, lj.epsilon, rc * unit.nanometer)
        PRNG.set_seed(1234)
        state = SamplerState(lj.positions, PRNG.get_random_key(), box_vectors=lj.box_vectors)
        ts = ThermodynamicState(pot_, temperature=140 * unit.kelvin, pressure=13.00765 * unit.atmosphere)
        nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer,
                               n_max_neighbors=400, builder="cell")
        nl.build_from_state(state)
        disp = MonteCarloDisplacementMove(displacement_sigma=0.0001 * unit.nanometer, number_of_moves=12)
        baro = MonteCarloBarostatMove(volume_max_scale=0.0001, number_of_moves=8)
        disp.device_loop = baro.device_loop = loop
        state, ts, nl = disp.update(state, ts, nl)
        state, ts, nl = baro.update(state, ts, nl)
        res[loop] = (disp.statistics, baro.statistics, _np(state.positions), _np(state.box_vectors),
                     np.asarray(state._current_PRNG_key).copy(), int(nl.n_neighbors.sum().item()))
    a, b = res[True], res[False]
    assert a[0] == b[0] and 0 < a[0]["n_accepted"] < 12
    assert a[1] == b[1] and a[1]["n_proposed"] == 8, (a[1], b[1])
    assert np.array_equal(a[4], b[4])
    assert np.allclose(a[3], b[3], rtol=1e-6) and np.allclose(a[2], b[2], rtol=2e-6, atol=2e-6)
    assert abs(a[5] - b[5]) <= 2      # list of the final configuration: positions equal to fp32 rounding


# ---------------------------------------------------------------------------------------------------
# generalisation: per-particle sigma / epsilon, energy shift
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shift,switch", [(False, 0.0), (True, 0.0), (False, 0.85)])
def test_lj_mixture_potential_vs_oracle(cuda_device, shift, switch):
    """Binary mixture with Lorentz-Berthelot mixing over a NeighborListNsqrd: energy and forces within rel 1e-5 of the
    float64 oracle; with uniform parameters and no shift it reproduces LJPotential; a Langevin run goes through the
    building-block path."""
    from chiron_b200 import unit
    from chiron_b200.integrators import LangevinIntegrator
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJMixturePotential, LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG
    lj_sys, x, box = _lj_system(9, 0.7, seed=17)
    n = x.shape[0]
    rng = np.random.default_rng(2)
    big = rng.random(n) < 0.3
    sig = np.where(big, 0.40, 0.34).astype(f32)
    eps = np.where(big, 1.6, 0.238 * 4.184).astype(f32)
    rc, skin = 1.02, 0.3
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer,
                           n_max_neighbors=200, builder="cell")
    nl.build(x, box)
    mix = LJMixturePotential(lj_sys.topology, sig * unit.nanometer, eps * unit.kilojoules_per_mole, rc * unit.nanometer,
                             shift=shift, switch_distance=(switch * unit.nanometer if switch else None))
    e, F = mix.compute_energy_and_force(x, nl)
    ref = pairs.build_neighborlist(x, box, rc, skin, int(nl.n_max_neighbors))
    e_ref, F_ref = pot.lj_mixture_energy_force_nlist(x.astype(np.float64), box.astype(np.float64), sig, eps, rc,
                                                     ref["neighbor_list"], ref["neighbor_mask"], shift=shift,
                                                     switch_distance=switch)
    assert np.isclose(float(e), e_ref, rtol=1e-5)
    assert np.allclose(_np(F), F_ref, rtol=1e-5, atol=1e-5 * np.abs(F_ref).max())
    assert abs(_np(F).astype(np.float64).sum(axis=0)).max() < 1e-4 * np.abs(F_ref).sum() / n      # Newton's third law
    if not shift and not switch:
        uni = LJMixturePotential(lj_sys.topology, np.full(n, 0.34, f32) * unit.nanometer,
                                 np.full(n, 0.238 * 4.184, f32) * unit.kilojoules_per_mole, rc * unit.nanometer)
        one = LJPotential(lj_sys.topology, 0.34 * unit.nanometer, 0.238 * unit.kilocalories_per_mole, rc * unit.nanometer)
        e1, F1 = uni.compute_energy_and_force(x, nl)
        e2, F2 = one.compute_energy_and_force(x, nl)
        assert np.isclose(float(e1), float(e2), rtol=1e-6) and np.allclose(_np(F1), _np(F2), rtol=1e-5, atol=1e-3)
    PRNG.set_seed(1234)
    state = SamplerState(positions=lj_sys.positions, current_PRNG_key=PRNG.get_random_key(), box_vectors=lj_sys.box_vectors)
    integ = LangevinIntegrator(timestep=1.0 * unit.femtosecond)
    out, _ = integ.run(state, ThermodynamicState(potential=mix, temperature=300 * unit.kelvin), number_of_steps=5, nbr_list=nl)
    assert integ.last_run_stats["path"] == "blocks" and np.isfinite(_np(out.positions)).all()
