"""MultiStateSampler on the GPU: the reference's own multistate tests (`chiron/tests/test_multistate.py`)
re-stated against this implementation, plus the batched / replica-exchange paths that are new here."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup_sampler(n_steps=100, **kw):
    from chiron_b200 import unit
    from chiron_b200.mcmc import LangevinDynamicsMove, MCMCSampler, MoveSchedule
    from chiron_b200.multistate import MultiStateSampler
    from chiron_b200.neighbors import OrthogonalNonPeriodicSpace, PairListNsqrd
    from chiron_b200.reporters import MultistateReporter
    nbr_list = PairListNsqrd(OrthogonalNonPeriodicSpace(), cutoff=10.0 * unit.nanometer)
    lang_move = LangevinDynamicsMove(timestep=1.0 * unit.femtoseconds, number_of_steps=n_steps)
    reporter = MultistateReporter()
    reporter.reset_reporter_file()
    sampler = MCMCSampler(MoveSchedule([("LangevinDynamicsMove", lang_move)]))
    return nbr_list, MultiStateSampler(mcmc_sampler=sampler, reporter=reporter, **kw)


def _ho_minima(cuda_device, tmp_path):
    """`ho_multistate_sampler_multiple_minima` (test_multistate.py:43-88)."""
    from chiron_b200 import unit
    from chiron_b200.potential import HarmonicOscillatorPotential
    from chiron_b200.reporters import BaseReporter
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import HarmonicOscillator
    from chiron_b200.utils import PRNG
    BaseReporter.set_directory(str(tmp_path))
    T = 300.0 * unit.kelvin
    x0s = [unit.Quantity(np.array([[x0, 0.0, 0.0]]), unit.angstrom) for x0 in np.linspace(0.0, 1.0, 3)]
    ho = HarmonicOscillator()
    thermo = [ThermodynamicState(HarmonicOscillatorPotential(ho.topology, x0=x0), temperature=T) for x0 in x0s]
    PRNG.set_seed(1234)
    states = [SamplerState(ho.positions, PRNG.get_random_key()) for _ in x0s]
    nbr_list, ms = _setup_sampler()
    ms.create(thermodynamic_states=thermo, sampler_states=states, nbr_lists=[copy.deepcopy(nbr_list) for _ in x0s])
    return ms


def test_multistate_class_and_minimize(cuda_device, tmp_path):
    ms = _ho_minima(cuda_device, tmp_path)
    assert ms._iteration == 0 and ms.number_of_replicas == 3 and ms.number_of_thermodynamic_states == 3
    assert ms._energy_thermodynamic_states.shape == (3, 3) and ms._n_proposed_matrix.shape == (3, 3)
    with pytest.raises(RuntimeError):
        ms.create(ms._thermodynamic_states, ms._sampler_states[:2], ms._nbr_lists)
    ms.minimize()
    x = [s.positions.cpu().numpy() for s in ms.sampler_states]
    assert np.allclose(x[0], [[0.0, 0.0, 0.0]], atol=1e-4)
    assert np.allclose(x[1], [[0.05, 0.0, 0.0]], atol=1e-2)       # test_multistate.py:196-205
    assert np.allclose(x[2], [[0.1, 0.0, 0.0]], atol=1e-2)


def test_multistate_run_free_energies(cuda_device, tmp_path):
    """`test_multistate_run` (test_multistate.py:211-251): 4 oscillators of different stiffness, 20 iterations,
    MBAR free energies within 0.1 of the analytic values."""
    from chiron_b200 import unit
    from chiron_b200.potential import HarmonicOscillatorPotential
    from chiron_b200.reporters import BaseReporter
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import HarmonicOscillator
    from chiron_b200.utils import PRNG
    BaseReporter.set_directory(str(tmp_path))
    ho = HarmonicOscillator()
    T = 300.0 * unit.kelvin
    kT = unit.BOLTZMANN_CONSTANT_kB * T * unit.AVOGADRO_CONSTANT_NA
    sigmas = [unit.Quantity(2.0 + 0.2 * k, unit.angstrom) for k in range(4)]
    thermo = [ThermodynamicState(HarmonicOscillatorPotential(ho.topology, k=kT / s ** 2), temperature=T) for s in sigmas]
    PRNG.set_seed(1234)
    states = [SamplerState(ho.positions, current_PRNG_key=PRNG.get_random_key()) for _ in sigmas]
    f_i = np.array([-np.log(2 * np.pi * float(s / unit.angstroms) ** 2) * 1.5 for s in sigmas])
    nbr_list, ms = _setup_sampler()
    ms.create(thermo, states, [copy.deepcopy(nbr_list) for _ in sigmas])
    n_iterations = 20
    ms.run(n_iterations)
    assert ms.iteration == n_iterations and ms.number_of_replicas == 4
    u_kn = ms._reporter.get_property("u_kn")
    assert u_kn.shape == (n_iterations + 1, 4, 4)                    # test_multistate.py:243
    assert np.allclose((f_i - f_i[0]), ms.f_k, atol=0.1)             # test_multistate.py:251


def _lj_replicas(n_rep, steps, tmp_path, batched, exchange=None, seed=3):
    from chiron_b200 import random as crandom, unit
    from chiron_b200.mcmc import LangevinDynamicsMove, MCMCSampler, MoveSchedule
    from chiron_b200.multistate import MultiStateSampler
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.reporters import BaseReporter, MultistateReporter
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import LennardJonesFluid
    BaseReporter.set_directory(str(tmp_path))
    lj = LennardJonesFluid(nparticles=512, reduced_density=0.8, seed=seed)
    pot = LJPotential(lj.topology, lj.sigma, lj.epsilon, 1.02 * unit.nanometer)
    temps = [300.0 * 1.15 ** k for k in range(n_rep)]
    thermo = [ThermodynamicState(pot, temperature=t * unit.kelvin) for t in temps]
    keys = crandom.split(crandom.PRNGKey(77), n_rep)
    states = [SamplerState(lj.positions, keys[k], box_vectors=lj.box_vectors) for k in range(n_rep)]
    nbrs = [NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=0.3 * unit.nanometer,
                              n_max_neighbors=200) for _ in range(n_rep)]
    move = LangevinDynamicsMove(timestep=1.0 * unit.femtoseconds, number_of_steps=steps)
    ms = MultiStateSampler(MCMCSampler(MoveSchedule([("LangevinDynamicsMove", move)])), MultistateReporter(),
                           exchange=exchange, exchange_seed=5, mcmc_iterations_per_sweep=1)
    ms.use_batched_engine = batched
    ms.create(thermo, states, nbrs)
    return ms, pot, temps


def test_batched_replicas_match_serial_path(cuda_device, tmp_path):
    """One engine launch for all replicas (blockIdx.y) == the reference's serial loop over replicas."""
    a, _, _ = _lj_replicas(4, 25, tmp_path, batched=True)
    b, _, _ = _lj_replicas(4, 25, tmp_path, batched=False)
    a.run(2)
    b.run(2)
    assert a._batched and not b._batched
    for sa, sb in zip(a.sampler_states, b.sampler_states):
        xa, xb = sa.positions.cpu().numpy(), sb.positions.cpu().numpy()
        assert np.allclose(xa, xb, atol=2e-5)
        assert np.allclose(sa.velocities.cpu().numpy(), sb.velocities.cpu().numpy(), atol=2e-4)
    assert np.allclose(a._energy_thermodynamic_states, b._energy_thermodynamic_states, rtol=1e-5)


def test_engine_groups_on_their_own_streams_match_one_engine(cuda_device, tmp_path, monkeypatch):
    """A handful of replicas per GPU run as two (or three) engines on separate streams and host threads
    (`_EngineGroups`): same swap history, positions and energies as the single engine."""
    from chiron_b200.multistate import _EngineGroups
    runs = {}
    for groups in ("1", "2", "3"):
        monkeypatch.setenv("CHX_REMD_ENGINE_GROUPS", groups)
        ms, _, _ = _lj_replicas(6, 30, tmp_path, batched=True, exchange="neighbors")
        ms.run(4)
        assert ms._batched
        assert isinstance(ms._batched.engine, _EngineGroups) == (groups != "1")
        if groups != "1":
            assert ms._batched.engine.G == int(groups)
            assert sorted(sum(ms._batched.engine.members, [])) == list(range(6))
        runs[groups] = (ms._replica_thermodynamic_states.copy(), np.array(ms._energy_thermodynamic_states),
                        [st.positions.cpu().numpy() for st in ms.sampler_states],
                        [st.velocities.cpu().numpy() for st in ms.sampler_states],
                        [np.asarray(st._current_PRNG_key).copy() for st in ms.sampler_states])
    monkeypatch.delenv("CHX_REMD_ENGINE_GROUPS")
    for groups in ("2", "3"):
        assert np.array_equal(runs["1"][0], runs[groups][0])
        assert np.allclose(runs["1"][1], runs[groups][1], rtol=1e-5)
        for xa, xb in zip(runs["1"][2], runs[groups][2]):
            assert np.allclose(xa, xb, atol=5e-5)
        for va, vb in zip(runs["1"][3], runs[groups][3]):
            assert np.allclose(va, vb, atol=5e-4)
        for ka, kb in zip(runs["1"][4], runs[groups][4]):
            assert np.array_equal(ka, kb)


@pytest.mark.parametrize("groups", ["1", "2"])
def test_prebuild_beside_the_exchange_phase_changes_nothing(cuda_device, tmp_path, monkeypatch, groups):
    """`chx_ljmd_set_prebuild`: the table rebuild a sweep's propagation ends on is enqueued on a side stream and
    the energies come from the cache filled before it -- same tables, so the trajectories, energies, swap history
    and keys are bit-identical to the run that rebuilds at the start of the next propagation."""
    monkeypatch.setenv("CHX_REMD_ENGINE_GROUPS", groups)
    runs = {}
    for pre in ("0", "1", "2"):
        monkeypatch.setenv("CHX_REMD_PREBUILD", pre)
        ms, _, _ = _lj_replicas(6, 100, tmp_path, batched=True, exchange="neighbors")
        ms.run(4)
        assert ms._batched
        runs[pre] = (ms._replica_thermodynamic_states.copy(), np.array(ms._energy_thermodynamic_states),
                     [st.positions.cpu().numpy() for st in ms.sampler_states],
                     [st.velocities.cpu().numpy() for st in ms.sampler_states],
                     [np.asarray(st._current_PRNG_key).copy() for st in ms.sampler_states],
                     ms._reporter.get_property("u_kn").copy())
    for mode in ("1", "2"):                              # 1: enqueued by the run, 2: behind the all-gather
        a, b = runs["0"], runs[mode]
        assert np.array_equal(a[0], b[0])
        assert np.allclose(a[1], b[1], rtol=1e-12)      # energies: fp64 atomics in any order
        for k in (2, 3, 4):
            for u, v in zip(a[k], b[k]):
                assert np.array_equal(u, v)
        assert np.allclose(a[5], b[5], rtol=1e-12)


def test_replica_exchange_energy_matrix_and_swaps(cuda_device, tmp_path):
    from chiron_b200 import unit
    ms, pot, temps = _lj_replicas(6, 20, tmp_path, batched=True, exchange="neighbors")
    ms.run(6)
    states = ms._replica_thermodynamic_states
    assert sorted(states.tolist()) == list(range(6))                 # one replica per state, always
    # u[r, k] = U_r / (R T_k): the matrix the swaps used equals a direct energy evaluation
    R = 8.314462618e-3
    for r, st in enumerate(ms.sampler_states):
        nl = ms._nbr_lists[r]
        nl.build(st.positions, st.box_vectors)
        U = float(pot.compute_energy(st.positions, nl))
        assert np.allclose(ms._energy_thermodynamic_states[r], [U / (R * t) for t in temps], rtol=2e-5)
    u = ms._reporter.get_property("u_kn")
    assert u.shape == (7, 6, 6)
    idx = ms._reporter.get_property("state_index")
    assert idx.shape == (7, 6) and any(not np.array_equal(idx[0], row) for row in idx[1:])   # something swapped


def test_minimization_like_reference(cuda_device):
    """`chiron/tests/test_minization.py`: 0 iterations leave the energy unchanged, 100 iterations lower it;
    two particles relax to r = 2^(1/6) sigma, E = -epsilon."""
    from chiron_b200 import unit
    from chiron_b200.minimze import minimize_energy
    from chiron_b200.neighbors import OrthogonalPeriodicSpace, PairListNsqrd
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState
    from chiron_b200.testsystems import LennardJonesFluid
    from chiron_b200.utils import PRNG
    lj = LennardJonesFluid(nparticles=216, reduced_density=0.1)
    cutoff = unit.Quantity(1.0, unit.nanometer)
    pot = LJPotential(lj.topology, cutoff=cutoff)
    PRNG.set_seed(1234)
    state = SamplerState(lj.positions, current_PRNG_key=PRNG.get_random_key(), box_vectors=lj.box_vectors)
    nbr = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=cutoff)
    nbr.build_from_state(state)
    e0 = float(pot.compute_energy(state.positions, nbr))
    e0_nolist = float(pot.compute_energy(state.positions))
    assert not np.isclose(e0, e0_nolist)                      # periodic images matter at this density
    r0 = minimize_energy(state.positions, pot.compute_energy, nbr, maxiter=0)
    assert np.isclose(float(pot.compute_energy(r0.params, nbr)), e0)
    r100 = minimize_energy(state.positions, pot.compute_energy, nbr, maxiter=100)
    e100 = float(pot.compute_energy(r100.params, nbr))
    assert e100 < e0 and not np.isnan(e100)
    # two particles
    sigma, eps = 1.0 * unit.nanometer, 1.0 * unit.kilojoules_per_mole
    pot2 = LJPotential(None, sigma=sigma, epsilon=eps, cutoff=3.0 * sigma)
    xy = np.array([[0.0, 0.0, 0.0], [0.9, 0.0, 0.0]], dtype=np.float32)
    st2 = SamplerState(positions=xy * unit.nanometer, current_PRNG_key=PRNG.get_random_key(),
                       box_vectors=np.eye(3, dtype=np.float32) * 10.0 * unit.nanometer)
    pl = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=3.0 * sigma)
    pl.build_from_state(st2)
    res = minimize_energy(st2.positions, pot2.compute_energy, pl, maxiter=10_000)
    x = res.params.cpu().numpy()
    assert np.isclose(float(pot2.compute_energy(res.params, pl)), -1.0, atol=1e-3)
    assert np.isclose(np.linalg.norm(x[1] - x[0]), 2 ** (1.0 / 6.0), atol=1e-3)
