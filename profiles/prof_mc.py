"""cProfile of the config-3 Monte Carlo moves (LJ NPT N = 32,768) to see where host time goes."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(n_side=32):
    import torch
    from loguru import logger
    logger.remove()
    from chiron_b200 import unit
    from chiron_b200.mcmc import MCMCSampler, MonteCarloBarostatMove, MonteCarloDisplacementMove, MoveSchedule
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import LennardJonesFluid
    from chiron_b200.utils import PRNG
    sigma, eps, rc, skin = 0.373, 0.2941, 1.4, 0.5
    n = n_side ** 3
    lj = LennardJonesFluid(nparticles=n, reduced_density=14.08 * sigma ** 3, sigma=sigma * unit.nanometer,
                           epsilon=eps * unit.kilocalories_per_mole, mass=16.04, seed=3, symbol="C")
    pot = LJPotential(lj.topology, lj.sigma, lj.epsilon, rc * unit.nanometer)
    PRNG.set_seed(1234)
    state = SamplerState(lj.positions, PRNG.get_random_key(), box_vectors=lj.box_vectors)
    thermo = ThermodynamicState(pot, temperature=140 * unit.kelvin, pressure=13.00765 * unit.atmosphere)
    nbr = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=rc * unit.nanometer, skin=skin * unit.nanometer,
                            n_max_neighbors=400, builder="cell")
    nbr.build_from_state(state)
    disp = MonteCarloDisplacementMove(displacement_sigma=0.0001 * unit.nanometer, number_of_moves=100,
                                      autotune=True, autotune_interval=100)
    baro = MonteCarloBarostatMove(volume_max_scale=0.0005, number_of_moves=10, autotune=True, autotune_interval=50)
    for name, move in (("displacement", disp), ("barostat", baro)):
        sampler = MCMCSampler(MoveSchedule([(name, move)]))
        state, thermo, nbr = sampler.run(state, thermo, 1, nbr)
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        state, thermo, nbr = sampler.run(state, thermo, 2, nbr)
        torch.cuda.synchronize()
        pr.disable()
        dt = time.perf_counter() - t0
        print(f"== {name}: {2 * move.number_of_moves / dt:.1f} moves/s, acceptance {move.n_accepted / move.n_proposed:.2f}")
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 32)
