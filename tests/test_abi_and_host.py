"""CPU-only checks: the C-ABI library builds, loads and exports every symbol the header declares; the
host-side key handling matches the oracle; the Python mirror of chiron's interface validates
inputs like the reference; and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "chiron_b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    lib = ctypes.CDLL(built_library)
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/chiron_b200.h but not exported"


def test_bindings_cover_header(built_library):
    from chiron_b200 import _lib, _engine  # noqa: F401  (registers the engine prototypes)
    declared = set(_declared_symbols()) - {"chx_last_error_string"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    _lib.load_library()


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "chiron_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_host_key_handling_matches_oracle(built_library):
    from chiron_b200 import random as crandom
    from chiron_b200.utils import PRNG
    from oracle import jax_random as jr
    from oracle import dynamics as dyn
    key = crandom.PRNGKey(1234)
    assert np.array_equal(key, jr.PRNGKey(1234))
    for _ in range(5):
        a, b = crandom.split(key), jr.split(key)
        assert np.array_equal(a, b)
        key = a[0]
    assert np.array_equal(crandom.split(key, 3), jr.split(key, 3))
    PRNG.set_seed(1234)
    stream = dyn.prng_stream(1234)
    for _ in range(3):
        k = PRNG.get_random_key()
        assert np.array_equal(k, next(stream))
        assert crandom.uniform_host(k) == jr.uniform(k)
        assert crandom.uniform_host(k, -1, 1) == jr.uniform(k, (), -1.0, 1.0)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built_library):
    from chiron_b200 import ChironB200Error, unit
    from chiron_b200.neighbors import OrthogonalPeriodicSpace, NeighborListNsqrd
    from chiron_b200.potential import LJPotential
    with pytest.raises(ChironB200Error):
        OrthogonalPeriodicSpace().wrap(np.zeros((2, 3), np.float32), np.eye(3, dtype=np.float32))
    with pytest.raises(ChironB200Error):
        LJPotential(None).compute_energy(np.zeros((2, 3), np.float32))
    # the raw ABI reports the missing device instead of computing anything
    lib = ctypes.CDLL(built_library)
    h = ctypes.c_void_p()
    assert lib.chx_context_create(0, None, ctypes.byref(h)) < 0
    lib.chx_last_error_string.restype = ctypes.c_char_p
    assert b"no CPU fallback" in lib.chx_last_error_string()


def test_unit_shim_matches_openmm_conventions():
    from chiron_b200 import unit
    kB = unit.BOLTZMANN_CONSTANT_kB * unit.AVOGADRO_CONSTANT_NA
    assert np.isclose((kB * 300 * unit.kelvin).value_in_unit_system(unit.md_unit_system), 2.494338785445972)
    assert (3.35 * unit.angstrom).value_in_unit_system(unit.md_unit_system) == 0.335
    assert np.isclose((1.0 * unit.kilocalories_per_mole).value_in_unit_system(unit.md_unit_system), 4.184)
    assert (2 * unit.femtosecond).value_in_unit_system(unit.md_unit_system) == 0.002
    assert (1.0 / unit.picosecond).value_in_unit_system(unit.md_unit_system) == 1.0
    assert (39.948 * unit.amu).value_in_unit_system(unit.md_unit_system) == 39.948
    beta = 1.0 / (unit.BOLTZMANN_CONSTANT_kB * 300 * unit.kelvin)
    pv = 1.0 * unit.atmosphere * (1000.0 * unit.nanometer ** 3)
    assert np.isclose(beta * pv, 24.4631, rtol=1e-5)
    assert not unit.Quantity(1.0, unit.radian).unit.is_compatible(unit.nanometer)
    assert unit.Quantity(1.0, unit.nanometer) == 10 * unit.angstrom
    q = 1.0 * unit.nanometer
    q *= 1.1
    assert np.isclose(q.value_in_unit(unit.nanometer), 1.1)


def test_state_and_list_validation_like_reference():
    """TypeError / ValueError contract of chiron/tests/test_states.py, test_pairs.py:163-237,
    test_potential.py:100-152, test_mcmc.py:289-324 (constructor-level checks need no GPU)."""
    from chiron_b200 import unit
    from chiron_b200.mcmc import (LangevinDynamicsMove, MCMCSampler, MonteCarloBarostatMove,
                                  MonteCarloDisplacementMove, MoveSchedule, MetropolisDisplacementMove)
    from chiron_b200.neighbors import (NeighborListNsqrd, OrthogonalNonPeriodicSpace,
                                       OrthogonalPeriodicSpace, PairListNsqrd)
    from chiron_b200.potential import HarmonicOscillatorPotential, IdealGasPotential, LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import HarmonicOscillator
    from chiron_b200.utils import PRNG

    PRNG.set_seed(1234)
    x = np.zeros((2, 3), np.float32)
    with pytest.raises(TypeError):
        SamplerState(x, PRNG.get_random_key())
    with pytest.raises(ValueError):
        SamplerState(unit.Quantity(x, unit.radian), PRNG.get_random_key())
    with pytest.raises(ValueError):
        SamplerState(unit.Quantity(x, unit.nanometer), None)
    with pytest.raises(ValueError):
        SamplerState(unit.Quantity(x, unit.nanometer), PRNG.get_random_key(),
                     box_vectors=unit.Quantity(np.zeros((4, 3)), unit.nanometer))
    with pytest.raises(ValueError):
        SamplerState(unit.Quantity(x, unit.nanometer), PRNG.get_random_key(),
                     velocities=unit.Quantity(np.zeros((3, 3)), unit.nanometer / unit.picosecond))
    with pytest.raises(TypeError):
        SamplerState(unit.Quantity(x, unit.nanometer), PRNG.get_random_key(), box_vectors=np.eye(3))
    state = SamplerState(unit.Quantity(x, unit.nanometer), PRNG.get_random_key())
    k0 = np.array(state._current_PRNG_key)
    sub = state.new_PRNG_key
    assert not np.array_equal(k0, state._current_PRNG_key) and sub.shape == (2,)

    space = OrthogonalPeriodicSpace()
    with pytest.raises(TypeError):
        NeighborListNsqrd(123, cutoff=1 * unit.nanometer, skin=0.1 * unit.nanometer)
    with pytest.raises(ValueError):
        NeighborListNsqrd(space, cutoff=unit.Quantity(1, unit.radian), skin=0.1 * unit.nanometer)
    with pytest.raises(ValueError):
        NeighborListNsqrd(space, cutoff=1 * unit.nanometer, skin=unit.Quantity(1, unit.radian))
    nl = NeighborListNsqrd(space, cutoff=1.1 * unit.nanometer, skin=0.1 * unit.nanometer, n_max_neighbors=5)
    assert nl.cutoff == 1.1 * unit.nanometer and nl.skin == 0.1 * unit.nanometer and nl.n_max_neighbors == 5
    assert nl.is_built is False
    with pytest.raises(TypeError):
        nl.build_from_state(123)
    with pytest.raises(TypeError):
        PairListNsqrd(123)
    with pytest.raises(ValueError):
        PairListNsqrd(space, cutoff=unit.Quantity(1, unit.radian))
    assert PairListNsqrd(OrthogonalNonPeriodicSpace(), cutoff=None).cutoff is None

    with pytest.raises(TypeError):
        LJPotential(123)
    with pytest.raises(TypeError):
        LJPotential(None, sigma=1.0)
    with pytest.raises(ValueError):
        LJPotential(None, sigma=1.0 * unit.kelvin)
    with pytest.raises(ValueError):
        LJPotential(None, epsilon=1.0 * unit.nanometer)
    with pytest.raises(ValueError):
        LJPotential(None, cutoff=1.0 * unit.kelvin)
    lj = LJPotential(None)
    assert lj.sigma == 0.335 and np.isclose(lj.epsilon, 4.184) and lj.cutoff == 1.0
    ho = HarmonicOscillator()
    with pytest.raises(TypeError):
        HarmonicOscillatorPotential(ho.topology, k=1.0)
    with pytest.raises(ValueError):
        HarmonicOscillatorPotential(ho.topology, k=1.0 * unit.nanometer)
    with pytest.raises(TypeError):
        IdealGasPotential(123)
    pot = HarmonicOscillatorPotential(ho.topology, ho.K, U0=ho.U0)
    assert np.isclose(pot.k, 41840.0)
    with pytest.raises(TypeError):
        ThermodynamicState(pot, temperature=300)
    with pytest.raises(ValueError):
        ThermodynamicState(pot, temperature=300 * unit.nanometer)
    with pytest.raises(ValueError):
        ThermodynamicState(pot, pressure=100 * unit.kelvin)
    with pytest.raises(ValueError):
        ThermodynamicState(pot, volume=1 * unit.kelvin)
    ts = ThermodynamicState(pot, temperature=300 * unit.kelvin, pressure=1 * unit.atmosphere)
    assert ts.nr_of_particles == 1 and ts.beta is not None

    with pytest.raises(ValueError):
        MoveSchedule([("bad", 123)])
    move = MonteCarloBarostatMove(volume_max_scale=0.1, number_of_moves=1)
    assert move.volume_max_scale == 0.1 and move.number_of_moves == 1
    disp = MonteCarloDisplacementMove(displacement_sigma=0.1 * unit.angstrom, number_of_moves=10)
    assert MetropolisDisplacementMove is MonteCarloDisplacementMove
    assert disp.statistics == dict(n_accepted=0, n_proposed=0)
    sampler = MCMCSampler(MoveSchedule([("d", disp), ("l", LangevinDynamicsMove(number_of_steps=3))]))
    assert len(sampler.move.move_schedule) == 2
    # autotune rules (mcmc.py:674-678, 907-911)
    disp.n_accepted, disp.n_proposed = 9, 10
    disp._autotune()
    assert np.isclose(disp.displacement_sigma.value_in_unit(unit.angstrom), 0.11)
    move.n_accepted, move.n_proposed = 9, 10
    move.volume_max_scale = 0.29
    move._autotune()
    assert move.volume_max_scale == 0.3
