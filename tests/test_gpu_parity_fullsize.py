"""GPU parity at BASELINE.json's own sizes (configs 1, 3, 4) against the CPU oracle, plus the cases the
small-N tests do not reach: an odd number of random words (3N odd) through the fused step kernel and the
LJ + NeighborListNsqrd barostat move (config 3's path).

The checker is `oracle.cport` (C/OpenMP restatement, cell-grid set-up accelerator proven equal to its
O(N^2) row builder in tests/test_oracle_c.py) where NumPy would take minutes, `oracle.dynamics`/`oracle.pairs`
elsewhere.  Bar: neighbour-list arrays bit-exact; positions within 1e-5 of the box edge, velocities within
1e-5 of their scale, keys bit-equal (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import cport
from oracle import dynamics as dyn
from oracle import pairs, potentials as pot

pytestmark = pytest.mark.gpu
f32 = np.float32
KT300 = 8.314462618e-3 * 300.0


def _np(t):
    return t.detach().cpu().numpy()


def _system(n_side, sigma, eps_kcal, rho_nm3, mass, seed):
    from chiron_b200 import unit
    from chiron_b200.testsystems import LennardJonesFluid
    lj = LennardJonesFluid(nparticles=n_side ** 3, reduced_density=rho_nm3 * sigma ** 3, sigma=sigma * unit.nanometer,
                           epsilon=eps_kcal * unit.kilocalories_per_mole, mass=mass, seed=seed)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=f32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=f32)
    return lj, x, box


# config 3: UA-TraPPE methane of Examples/LJ_MCMC.py (sigma 0.373 nm, rc 1.4 nm, 14.08 / nm^3), N = 32,768
# config 4: LJ argon, rho* = 0.8, rc = 3 sigma, N = 262,144
_CONFIGS = {
    "config3_n32768": dict(n_side=32, sigma=0.373, eps=0.2941, rho=14.08, mass=16.04, rc=1.4, skin=0.5, M=400),
    "config4_n262144": dict(n_side=64, sigma=0.34, eps=0.238, rho=0.8 / 0.34 ** 3, mass=39.948, rc=1.02, skin=0.5,
                            M=320),
}


@pytest.mark.parametrize("name", list(_CONFIGS))
def test_cell_list_arrays_equal_the_oracle_at_baseline_sizes(cuda_device, name):
    """`NeighborListNsqrd(builder="cell").build` (neighbors.py:595-626, 671-729 semantics): n_neighbors,
    neighbor_list and the pad mask ARRAY-EQUAL to the C oracle's list, also after the particles have moved
    off the lattice and partly out of the box (the reference builds from unwrapped positions)."""
    from chiron_b200 import unit
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    c = _CONFIGS[name]
    cport.use_all_cores()
    lj, x, box = _system(c["n_side"], c["sigma"], c["eps"], c["rho"], c["mass"], seed=4)
    rng = np.random.default_rng(17)
    for trial in range(2):
        if trial == 1:   # thermal-size displacements + a rigid shift that pushes a slab out of the box
            # (0.02 nm: no near-contact pairs, whose (sigma/d)^12 would amplify the rounding of the
            # reference's own `r + L/2` step far beyond 1e-5)
            x = (x + rng.normal(size=x.shape).astype(f32) * f32(0.02) + f32(0.37)).astype(f32)
        nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=c["rc"] * unit.nanometer,
                               skin=c["skin"] * unit.nanometer, n_max_neighbors=c["M"], builder="cell")
        nl.build(x, box)
        M = int(nl.n_max_neighbors)
        ol, om, on, omax = cport.build_cells(x, box, f32(c["rc"] + c["skin"]), M)
        assert omax < M
        assert np.array_equal(_np(nl.n_neighbors).astype(np.int32), on)
        assert np.array_equal(_np(nl.neighbor_list).astype(np.uint32), ol)
        assert np.array_equal(_np(nl.neighbor_mask).astype(np.int32), om)
        # calculate(): interacting pairs of the list == the oracle's energy loop count, energy within 1e-5
        from chiron_b200.potential import LJPotential
        potential = LJPotential(lj.topology, lj.sigma, lj.epsilon, c["rc"] * unit.nanometer)
        e_gpu = float(potential.compute_energy(x, nl))
        e_ref, _, n_int = cport.lj_nlist(x, box, c["sigma"], c["eps"] * 4.184, c["rc"], ol, om, want_force=False)
        _, _, mask, _, _ = nl.calculate(torch.as_tensor(x, device=cuda_device))
        assert int(mask.sum().item()) == n_int
        assert np.isclose(e_gpu, e_ref, rtol=1e-5)
        del nl, mask, ol, om


def _langevin_vs_cport(cuda_device, n_side, nsteps, skin, M, seed, dt_fs=1.0, rho_star=0.8):
    from chiron_b200 import unit
    from chiron_b200.integrators import LangevinIntegrator
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG
    cport.use_all_cores()
    sigma, eps, rc = 0.34, 0.238 * 4.184, 1.02
    lj, x, box = _system(n_side, 0.34, 0.238, rho_star / 0.34 ** 3, 39.948, seed=seed)
    n = x.shape[0]
    mass = np.full(n, 39.948, f32)
    PRNG.set_seed(1234)
    state = SamplerState(positions=lj.positions, current_PRNG_key=PRNG.get_random_key(), box_vectors=lj.box_vectors)
    v0 = dyn.maxwell_boltzmann(np.array([7, 11], np.uint32), mass, 300.0)
    state.velocities = torch.as_tensor(v0, device=cuda_device)
    potential = LJPotential(lj.topology, lj.sigma, lj.epsilon, rc * unit.nanometer)
    ts = ThermodynamicState(potential=potential, temperature=300 * unit.kelvin)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer,
                           n_max_neighbors=M, builder="cell")
    integ = LangevinIntegrator(timestep=dt_fs * unit.femtosecond)
    out, nl_out = integ.run(state, ts, number_of_steps=nsteps, nbr_list=nl)
    assert integ.last_run_stats["path"] == "fused"
    # oracle: same loop key (SamplerState.new_PRNG_key, states.py:150-154), same velocities
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))
    key = st.new_key()
    cport.set_build_mode(1)
    try:
        xo, vo, key_out, stats = cport.langevin_lj(x, v0, mass, box, sigma, eps, rc, skin, M, KT300, dt_fs * 1e-3, 1.0,
                                                   key, nsteps)
    finally:
        cport.set_build_mode(0)
    L = np.diag(box)
    dx = _np(out.positions) - xo
    dx -= L * np.round(dx / L)
    return dict(dx=np.abs(dx).max(), L=float(L.max()), vg=_np(out.velocities), vo=vo, key_gpu=np.asarray(out.current_PRNG_key),
                key_o=key_out, stats=stats, run=integ.last_run_stats, nl=nl_out)


def test_fused_engine_five_steps_vs_oracle_at_262144(cuda_device):
    """BASELINE config 4 at full size: 5 BAOAB steps of the fused engine from the reference's jax.random
    stream against `cport.langevin_lj` (integrators.py:174-195) -- positions within 1e-5 L, velocities within
    1e-5 of their scale, loop key bit-equal."""
    r = _langevin_vs_cport(cuda_device, 64, 5, 0.5, 320, seed=4)
    assert r["dx"] < 1e-5 * r["L"], r["dx"]
    assert np.allclose(r["vg"], r["vo"], rtol=1e-5, atol=1e-5 * float(np.abs(r["vo"]).max()))
    assert np.array_equal(r["key_gpu"], r["key_o"])
    assert r["stats"]["p_int"] == r["run"]["interacting_pairs"] or r["run"]["interacting_pairs"] == 0


def test_config1_lj_langevin_1000_particles_vs_oracle(cuda_device):
    """BASELINE config 1 (Examples/LJ_langevin.py:6-90: N = 1000, rho* = 0.1, skin 0.5 nm, n_max_neighbors = 180):
    25 steps through LangevinIntegrator.run against the oracle.  (At this density no row comes near 180
    entries, so the reference's silent truncation of rows longer than n_max_neighbors -- SURVEY App. B #1,
    which the oracle replicates and the engine does not -- plays no role.)"""
    r = _langevin_vs_cport(cuda_device, 10, 25, 0.5, 180, seed=6, rho_star=0.1)
    assert r["stats"]["M"] == 180
    assert r["dx"] < 2e-5 * r["L"], r["dx"]
    assert np.allclose(r["vg"], r["vo"], rtol=1e-4, atol=2e-5 * float(np.abs(r["vo"]).max()))
    assert np.array_equal(r["key_gpu"], r["key_o"])


def test_fused_engine_odd_number_of_random_words(cuda_device):
    """7^3 = 343 particles: 3N = 1029 is odd, so jax.random's counter array is padded with one zero and the
    halves of the threefry output are split at ceil(3N/2) -- through k_md_force<UPDATE>
    against the NumPy oracle, 1 and 8 steps."""
    from chiron_b200 import unit
    from chiron_b200.integrators import LangevinIntegrator
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG
    sigma, eps, rc, skin = 0.34, 0.238 * 4.184, 1.02, 0.2
    for nsteps in (1, 8):
        lj, x, box = _system(7, 0.34, 0.238, 0.8 / 0.34 ** 3, 39.948, seed=23)
        assert x.shape[0] == 343 and (3 * x.shape[0]) % 2 == 1
        PRNG.set_seed(1234)
        state = SamplerState(positions=lj.positions, current_PRNG_key=PRNG.get_random_key(),
                             box_vectors=lj.box_vectors)
        potential = LJPotential(lj.topology, lj.sigma, lj.epsilon, rc * unit.nanometer)
        ts = ThermodynamicState(potential=potential, temperature=300 * unit.kelvin)
        nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer,
                               n_max_neighbors=180, builder="cell")
        integ = LangevinIntegrator(timestep=1.0 * unit.femtosecond)
        out, _ = integ.run(state, ts, number_of_steps=nsteps, nbr_list=nl)   # velocities drawn from the key
        assert integ.last_run_stats["path"] == "fused"
        nbr = dyn.OracleNeighborList(box, rc, skin, 180)
        st = dyn.KeyedState(next(dyn.prng_stream(1234)))
        force = lambda xx: pot.lj_force_nlist(xx, box, sigma, eps, rc, nbr.neighbor_list, nbr.neighbor_mask)  # noqa: E731
        xo, vo, key, _ = dyn.langevin_run(x, None, np.full(343, 39.948), 300.0, 0.001, 1.0, st, nsteps, force, nbr=nbr)
        L = np.diag(box)
        dx = _np(out.positions) - xo
        dx -= L * np.round(dx / L)
        assert np.abs(dx).max() < 1e-5 * float(L.max())
        assert np.allclose(_np(out.velocities), vo, rtol=1e-5, atol=1e-5 * float(np.abs(vo).max()))
        assert np.array_equal(np.asarray(out.current_PRNG_key), key)


@pytest.mark.parametrize("device_loop", [True, False])
def test_mc_barostat_lj_neighborlist_vs_oracle(cuda_device, device_loop):
    """MonteCarloBarostatMove on an LJ fluid with a NeighborListNsqrd (BASELINE config 3's path,
    mcmc.py:913-1009): accept sequence, box, positions and PRNG key follow `dyn.mc_barostat_step(..., nbr=...)`
    over 24 moves, through the device-resident loop and through the building blocks."""
    from chiron_b200 import unit
    from chiron_b200.mcmc import MonteCarloBarostatMove
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG
    sigma, eps, rc, skin, P_atm, n_moves, scale = 0.34, 0.238 * 4.184, 1.02, 0.3, 2000.0, 24, 0.002
    lj, x, box = _system(12, 0.34, 0.238, 0.8 / 0.34 ** 3, 39.948, seed=91)
    potential = LJPotential(lj.topology, lj.sigma, lj.epsilon, rc * unit.nanometer)
    PRNG.set_seed(1234)
    state = SamplerState(positions=lj.positions, current_PRNG_key=PRNG.get_random_key(), box_vectors=lj.box_vectors)
    ts = ThermodynamicState(potential=potential, temperature=300 * unit.kelvin, pressure=P_atm * unit.atmosphere)
    nl = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer,
                           n_max_neighbors=180, builder="cell")
    nl.build_from_state(state)
    move = MonteCarloBarostatMove(volume_max_scale=scale, number_of_moves=n_moves)
    move.device_loop = device_loop
    out, ts_out, nl_out = move.update(state, ts, nl)

    nbr = dyn.OracleNeighborList(box, rc, skin, 180)
    nbr.build(x)
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))

    def red(xx, bb):
        V = f32(f32(bb[0, 0] * bb[1, 1]) * bb[2, 2])
        U = pot.lj_energy_nlist(xx, bb, sigma, eps, rc, nbr.neighbor_list, nbr.neighbor_mask)
        return dyn.reduced_potential(U, 300.0, P_atm, V)
    xo, bo, u, accepted = x, box, red(x, box), []
    for _ in range(n_moves):
        xo, bo, u, a = dyn.mc_barostat_step(xo, bo, st, scale, u, red, nbr=nbr)
        accepted.append(a)
    assert 0 < sum(accepted) < n_moves
    assert move.statistics == dict(n_accepted=int(sum(accepted)), n_proposed=n_moves)
    assert np.array_equal(np.asarray(out._current_PRNG_key), st.key)
    assert np.allclose(_np(out.box_vectors), bo, rtol=2e-6)
    assert np.allclose(_np(out.positions), xo, rtol=4e-6, atol=2e-6)
    # the list handed back: the reference list of the final configuration, pad mask included
    xg, bg = _np(out.positions), _np(out.box_vectors)
    ref = pairs.build_neighborlist(xg, bg, rc, skin, int(nl_out.n_max_neighbors))
    assert np.array_equal(_np(nl_out.n_neighbors), ref["n_neighbors"])
    assert np.array_equal(_np(nl_out.neighbor_list).astype(np.uint32), ref["neighbor_list"])
    assert np.array_equal(_np(nl_out.neighbor_mask).astype(np.int32), ref["neighbor_mask"].astype(np.int32))
    assert abs(int(ref["n_neighbors"].sum()) - int(nbr.n_neighbors.sum())) <= 2
