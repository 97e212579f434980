"""cProfile of the end-to-end call bench.py times (LangevinIntegrator.run with host buffers, 100 steps)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from loguru import logger
    logger.remove()
    from chiron_b200 import random as crandom, unit
    from chiron_b200.integrators import LangevinIntegrator
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG, initialize_velocities
    dev = torch.device("cuda", 0)
    lj, x, box = bench.make_system(64, seed=4)
    n = x.shape[0]
    v0 = initialize_velocities(bench.TEMP_K * unit.kelvin, lj.topology, crandom.PRNGKey(11))
    v0 = v0.value_in_unit_system(unit.md_unit_system).cpu().numpy()
    potential = LJPotential(lj.topology, lj.sigma, lj.epsilon, bench.RC * unit.nanometer)
    nbr = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=bench.RC * unit.nanometer, skin=bench.SKIN * unit.nanometer,
                            n_max_neighbors=400, builder="cell")
    ts = ThermodynamicState(potential, temperature=bench.TEMP_K * unit.kelvin)
    integ = LangevinIntegrator(timestep=bench.DT_PS * unit.picosecond, collision_rate=bench.GAMMA / unit.picosecond)
    hx, hv = torch.from_numpy(x).pin_memory(), torch.from_numpy(v0).pin_memory()
    ox, ov = torch.empty((n, 3)).pin_memory(), torch.empty((n, 3)).pin_memory()
    PRNG.set_seed(1234)
    key = PRNG.get_random_key()

    def one_call(key):
        state = SamplerState(unit.Quantity(hx.to(dev, non_blocking=True), unit.nanometer), key,
                             velocities=unit.Quantity(hv.to(dev, non_blocking=True), unit.nanometer / unit.picosecond),
                             box_vectors=unit.Quantity(box, unit.nanometer))
        out, _ = integ.run(state, ts, number_of_steps=100, nbr_list=nbr)
        ox.copy_(out.positions, non_blocking=True)
        ov.copy_(out.velocities, non_blocking=True)
        energy = float(integ._engine.energy()[0])
        torch.cuda.synchronize()
        return out._current_PRNG_key, energy
    for _ in range(3):
        key, _ = one_call(key)
    # wall time per engine entry point (each of them ends in a stream synchronisation or is tiny)
    from chiron_b200 import _engine
    acc = {}
    orig = _engine.LJLangevinEngine._call

    def timed(self, name, *a):
        torch.cuda.synchronize()
        t = time.perf_counter()
        orig(self, name, *a)
        torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
    _engine.LJLangevinEngine._call = timed
    for _ in range(5):
        key, _ = one_call(key)
    print({k: round(v / 5 * 1e3, 3) for k, v in acc.items()}, "ms per call")
    _engine.LJLangevinEngine._call = orig
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for _ in range(10):
        key, _ = one_call(key)
    pr.disable()
    print("ms per call", (time.perf_counter() - t0) * 100)
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)


if __name__ == "__main__":
    main()
