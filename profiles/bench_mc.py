"""Monte Carlo configs of BASELINE.json on the GPU (secondary numbers, recorded under profiles/):
  config 3: LJ NPT N = 32,768 (UA-TraPPE methane parameters of Examples/LJ_MCMC.py): 100 all-particle
            displacement moves + 10 barostat moves per iteration (full energy re-evaluation, list rebuild
            on every volume move, like the reference);
  config 2: single-particle displacement on LJ N = 1000 through the subset delta-energy kernel vs the
            reference's full re-evaluation.
    python profiles/bench_mc.py [n_side]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(n_side=32, quiet=False):
    import torch
    from chiron_b200 import unit
    from chiron_b200.mcmc import MCMCSampler, MonteCarloBarostatMove, MonteCarloDisplacementMove, MoveSchedule
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import LennardJonesFluid
    from chiron_b200.utils import PRNG
    from loguru import logger
    if quiet:
        logger.remove()
    out = {"api": "MCMCSampler.run (device-resident Metropolis loops: chx_mc_displace_run, chx_mc_barostat_run)"}
    # ---- config 3 ----
    sigma, eps, rc, skin = 0.373, 0.2941, 1.4, 0.5
    n = n_side ** 3
    rho_star = 14.08 * sigma ** 3
    lj = LennardJonesFluid(nparticles=n, reduced_density=rho_star, sigma=sigma * unit.nanometer,
                           epsilon=eps * unit.kilocalories_per_mole, mass=16.04, seed=3, symbol="C")
    pot = LJPotential(lj.topology, lj.sigma, lj.epsilon, rc * unit.nanometer)
    PRNG.set_seed(1234)
    state = SamplerState(lj.positions, PRNG.get_random_key(), box_vectors=lj.box_vectors)
    thermo = ThermodynamicState(pot, temperature=140 * unit.kelvin, pressure=13.00765 * unit.atmosphere)
    nbr = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer,
                            n_max_neighbors=400, builder="cell")
    nbr.build_from_state(state)
    # SURVEY.md section 8d names sigma_disp = 0.001 nm / volume scale 0.1 with autotune: at N = 32,768 those
    # start at 0 % acceptance, so the timed moves start from the values autotune converges to
    disp = MonteCarloDisplacementMove(displacement_sigma=0.0001 * unit.nanometer, number_of_moves=100,
                                      autotune=True, autotune_interval=100)
    baro = MonteCarloBarostatMove(volume_max_scale=0.0005, number_of_moves=10, autotune=True, autotune_interval=50)
    for name, move, reps in (("displacement", disp, 5), ("barostat", baro, 5)):
        sampler = MCMCSampler(MoveSchedule([(name, move)]))
        state, thermo, nbr = sampler.run(state, thermo, 1, nbr)      # warm-up (allocations, autotune)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            state, thermo, nbr = sampler.run(state, thermo, 1, nbr)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        moves = reps * move.number_of_moves
        out[f"cfg3_{name}_moves_per_s"] = moves / dt
        out[f"cfg3_{name}_acceptance"] = move.n_accepted / max(1, move.n_proposed)
    out["cfg3"] = {"n": n, "box_nm": float(state.box_lengths_host()[0]), "n_max_neighbors": int(nbr.n_max_neighbors)}
    # ---- config 2: single-particle moves, delta-energy kernel vs full re-evaluation ----
    lj = LennardJonesFluid(nparticles=1000, reduced_density=0.8, seed=1)
    pot = LJPotential(lj.topology, lj.sigma, lj.epsilon, 1.02 * unit.nanometer)
    thermo = ThermodynamicState(pot, temperature=300 * unit.kelvin)
    for label, use_delta in (("delta_kernel", True), ("full_energy", False)):
        PRNG.set_seed(1234)
        state = SamplerState(lj.positions, PRNG.get_random_key(), box_vectors=lj.box_vectors)
        nbr = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=0.5 * unit.nanometer,
                                n_max_neighbors=400)
        nbr.build_from_state(state)
        move = MonteCarloDisplacementMove(displacement_sigma=0.01 * unit.nanometer, number_of_moves=500,
                                          atom_subset=[17], use_delta_energy=use_delta)
        sampler = MCMCSampler(MoveSchedule([("d", move)]))
        state, thermo, nbr = sampler.run(state, thermo, 1, nbr)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        state, thermo, nbr = sampler.run(state, thermo, 2, nbr)
        torch.cuda.synchronize()
        out[f"cfg2_single_particle_{label}_moves_per_s"] = 1000 / (time.perf_counter() - t0)
        out[f"cfg2_single_particle_{label}_acceptance"] = move.n_accepted / max(1, move.n_proposed)
    # ---- config 2: harmonic oscillator (N = 5, tests/test_mcmc.py:166-286) and ideal gas (N = 216, Examples/Idealgas.py) ----
    from chiron_b200.neighbors import PairListNsqrd
    from chiron_b200.potential import HarmonicOscillatorPotential, IdealGasPotential
    from chiron_b200.testsystems import _topology
    ho = HarmonicOscillatorPotential(_topology(5), k=100.0 * unit.kilocalories_per_mole / unit.angstrom ** 2,
                                     x0=np.zeros((5, 3)) * unit.angstrom)
    PRNG.set_seed(1234)
    state = SamplerState(np.zeros((5, 3), np.float32) * unit.nanometer, PRNG.get_random_key())
    thermo = ThermodynamicState(ho, temperature=300 * unit.kelvin)
    move = MonteCarloDisplacementMove(displacement_sigma=0.1 * unit.angstrom, number_of_moves=1000)
    sampler = MCMCSampler(MoveSchedule([("d", move)]))
    state, thermo, _ = sampler.run(state, thermo, 1, None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    state, thermo, _ = sampler.run(state, thermo, 5, None)
    torch.cuda.synchronize()
    out["cfg2_harmonic_oscillator_moves_per_s"] = 5000 / (time.perf_counter() - t0)
    out["cfg2_harmonic_oscillator_acceptance"] = move.n_accepted / max(1, move.n_proposed)
    rng = np.random.default_rng(3)
    box = np.eye(3, dtype=np.float32) * 20.6
    PRNG.set_seed(1234)
    state = SamplerState((rng.random((216, 3)) * 20.6).astype(np.float32) * unit.nanometer, PRNG.get_random_key(),
                         box_vectors=box * unit.nanometer)
    thermo = ThermodynamicState(IdealGasPotential(_topology(216)), temperature=298 * unit.kelvin,
                                pressure=1.0 * unit.atmosphere)
    nbr = PairListNsqrd(OrthogonalPeriodicSpace(), cutoff=0 * unit.nanometer)
    nbr.build_from_state(state)
    move = MonteCarloDisplacementMove(displacement_sigma=0.1 * unit.nanometer, number_of_moves=1000)
    sampler = MCMCSampler(MoveSchedule([("d", move)]))
    state, thermo, nbr = sampler.run(state, thermo, 1, nbr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    state, thermo, nbr = sampler.run(state, thermo, 5, nbr)
    torch.cuda.synchronize()
    out["cfg2_ideal_gas_moves_per_s"] = 5000 / (time.perf_counter() - t0)
    if not quiet:
        print(json.dumps(out))
    return out


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32)
