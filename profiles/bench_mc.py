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


def main(n_side=32, quiet=False, cpu_baselines=False):
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
    # ---- config 1: Examples/LJ_langevin.py:6-90 (N = 1000, rho* = 0.1, skin 0.5 nm, n_max_neighbors = 180, 1000 steps) ----
    from chiron_b200.integrators import LangevinIntegrator
    lj = LennardJonesFluid(nparticles=1000, reduced_density=0.1, seed=1)
    pot = LJPotential(lj.topology, lj.sigma, lj.epsilon, 1.02 * unit.nanometer)
    thermo = ThermodynamicState(pot, temperature=300 * unit.kelvin)
    PRNG.set_seed(1234)
    state = SamplerState(lj.positions, PRNG.get_random_key(), box_vectors=lj.box_vectors)
    nbr = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=1.02 * unit.nanometer, skin=0.5 * unit.nanometer,
                            n_max_neighbors=180)
    nbr.build_from_state(state)
    integ = LangevinIntegrator(timestep=1.0 * unit.femtosecond, report_interval=100)
    state, nbr = integ.run(state, thermo, number_of_steps=1000, nbr_list=nbr)       # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        state, nbr = integ.run(state, thermo, number_of_steps=1000, nbr_list=nbr)
    torch.cuda.synchronize()
    out["cfg1_lj_langevin_n1000_steps_per_s"] = 3000 / (time.perf_counter() - t0)
    out["cfg1"] = {"api": "LangevinIntegrator.run(number_of_steps=1000)", "path": integ.last_run_stats.get("path"),
                   "n": 1000, "rho_star": 0.1}
    # ---- launch-latency floor for the latency-bound config 2 numbers: one empty-ish kernel per graph node ----
    one = torch.zeros(32, device="cuda")
    g = torch.cuda.CUDAGraph()
    s_ = torch.cuda.Stream()
    with torch.cuda.stream(s_):
        one.add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s_):
            for _ in range(200):
                one.add_(1.0)
    g.replay(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    node_us = (time.perf_counter() - t0) / 2000 * 1e6
    out["latency_floor"] = {"graph_kernel_node_us": node_us,
                            "moves_per_s_at_3_nodes_per_move": 1e6 / (3 * node_us),
                            "note": "a Metropolis move of the graph loop is 3 dependent launches (propose, energy, decide); "
                                    "k_mcl_small runs all moves of small systems inside ONE launch and is not bound by this"}
    if cpu_baselines:
        out["cpu_baseline"] = cpu_baselines_mc(n_side)
    if not quiet:
        print(json.dumps(out))
    return out


def cpu_baselines_mc(n_side=32):
    """The reference algorithm on the host cores for configs 1, 2, 3 (C/OpenMP port where one exists, NumPy oracle
    for the tiny systems); bounded samples, a few seconds in total."""
    from oracle import cport, dynamics as dyn, jax_random as jr, pairs, potentials as opot
    from chiron_b200 import unit
    from chiron_b200.testsystems import LennardJonesFluid
    cport.use_all_cores()
    f32 = np.float32
    res = {"cores": cport.num_threads(), "kind": "port"}
    # config 1: the whole run
    lj = LennardJonesFluid(nparticles=1000, reduced_density=0.1, seed=1)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=f32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=f32)
    v0 = dyn.maxwell_boltzmann(jr.PRNGKey(11), np.full(1000, 39.948), 300.0)
    t0 = time.perf_counter()
    cport.langevin_lj(x, v0, np.full(1000, 39.948, f32), box, 0.34, 0.238 * 4.184, 1.02, 0.5, 180, 8.314462618e-3 * 300,
                      0.001, 1.0, np.array([0, 1], np.uint32), 1000)
    res["cfg1_lj_langevin_n1000_steps_per_s"] = 1000 / (time.perf_counter() - t0)
    res["cfg1_sample"] = "the full 1000-step run incl. O(N^2) builds, C/OpenMP port"
    # config 3: one displacement move = noise for 3N coordinates + wrap + check + full LJ energy over the half list;
    # one barostat move = rescale + the reference's O(N^2) list build + full energy
    sigma, eps, rc, skin = 0.373, 0.2941 * 4.184, 1.4, 0.5
    n = n_side ** 3
    lj = LennardJonesFluid(nparticles=n, reduced_density=14.08 * sigma ** 3, sigma=sigma * unit.nanometer,
                           epsilon=0.2941 * unit.kilocalories_per_mole, mass=16.04, seed=3, symbol="C")
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=f32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=f32)
    nl, mask, nn, mx = cport.build_cells(x, box, f32(rc + skin), 440)
    key = np.array([0, 7], np.uint32)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        xi = cport.normal(key, (n, 3))
        xp = cport.wrap((x + xi * f32(1e-4)).astype(f32), box)
        cport.check(xp, x, box, skin)
        cport.lj_nlist(xp, box, sigma, eps, rc, nl, mask, want_force=False)
    res["cfg3_displacement_moves_per_s"] = reps / (time.perf_counter() - t0)
    t0 = time.perf_counter()
    xp = (x * f32(1.0001)).astype(f32)
    bp = (box * f32(1.0001)).astype(f32)
    nl2, mask2, nn2, _ = cport.build_rows(xp, bp, f32(rc + skin), 440)
    cport.lj_nlist(xp, bp, sigma, eps, rc, nl2, mask2, want_force=False)
    res["cfg3_barostat_moves_per_s"] = 1.0 / (time.perf_counter() - t0)
    res["cfg3_sample"] = "%d displacement moves and 1 barostat move at N=%d executed in full (C/OpenMP port)" % (reps, n)
    # config 2: single-particle move on LJ N = 1000 (the reference re-evaluates the full energy), NumPy-free C calls
    lj = LennardJonesFluid(nparticles=1000, reduced_density=0.8, seed=1)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=f32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=f32)
    nl, mask, nn, mx = cport.build_rows(x, box, f32(1.52), 400)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        xi = cport.normal(key, (1000, 3))
        xp = cport.wrap((x + xi * f32(1e-2)).astype(f32), box)
        cport.check(xp, x, box, 0.5)
        cport.lj_nlist(xp, box, 0.34, 0.238 * 4.184, 1.02, nl, mask, want_force=False)
    res["cfg2_single_particle_full_energy_moves_per_s"] = reps / (time.perf_counter() - t0)
    # harmonic oscillator N = 5 / ideal gas N = 216 through the NumPy oracle's step function
    st = dyn.KeyedState(next(dyn.prng_stream(1234)))
    x5 = np.zeros((5, 3), f32)
    k_ho = 100.0 * 4.184 * 100.0
    red = lambda xx, b: dyn.reduced_potential(opot.ho_energy(xx, np.zeros((5, 3), f32), k_ho, 0.0), 300.0)  # noqa: E731
    try:
        u = red(x5, None)
        t0 = time.perf_counter()
        for _ in range(300):
            x5, u, _a = dyn.mc_displacement_step(x5, None, st, 0.01, u, red)
        res["cfg2_harmonic_oscillator_moves_per_s"] = 300 / (time.perf_counter() - t0)
    except Exception as exc:
        res["cfg2_harmonic_oscillator_moves_per_s"] = None
        res["cfg2_harmonic_oscillator_error"] = repr(exc)
    res["cfg2_sample"] = "200 single-particle LJ moves (C port), 300 harmonic-oscillator moves (NumPy oracle, interpreter bound)"
    return res


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32)
