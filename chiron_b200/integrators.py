"""LangevinIntegrator (BAOAB) with chiron's interface (`chiron/integrators.py`).

Two execution paths, same numbers:
  * fused engine (`chx_ljmd_*`): LJPotential + NeighborListNsqrd under periodic boundaries -- the
    whole trajectory runs on the device (cell-sorted particles, tiled neighbour structure, fused
    BAOAB+wrap+check kernel, device-side rebuild decision);
  * building blocks (`chx_baoab_update` + `potential.compute_force`): any other potential / list.
Both consume the reference's `jax.random` stream, so a trajectory started from the same
SamplerState key reproduces the JAX path step by step.
"""
import copy
import math
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib, random, unit
from .states import SamplerState, ThermodynamicState


class LangevinIntegrator:
    def __init__(self, timestep=1.0 * unit.femtoseconds, collision_rate=1.0 / unit.picoseconds,
                 refresh_velocities: bool = False, report_interval: int = 100, reporter=None,
                 save_traj_in_memory: bool = False) -> None:
        from loguru import logger as log
        self.kB = unit.BOLTZMANN_CONSTANT_kB * unit.AVOGADRO_CONSTANT_NA
        log.info(f"timestep = {timestep}")
        log.info(f"collision_rate = {collision_rate}")
        log.info(f"report_interval = {report_interval}")
        self.timestep = timestep
        self.collision_rate = collision_rate
        if reporter:
            self.reporter = reporter
        self.report_interval = report_interval
        self.velocities = None
        self.save_traj_in_memory = save_traj_in_memory
        self.traj = []
        self.refresh_velocities = refresh_velocities
        self._move_iteration = 0
        self.use_fused_engine = True   # set False to force the building-block path
        self.last_run_stats = {}

    # -- constants of the scheme in the reference's precision (integrators.py:127-137) --------------
    def _coefficients(self, temperature):
        kT = float((self.kB * temperature).value_in_unit_system(unit.md_unit_system))
        dt = float(self.timestep.value_in_unit_system(unit.md_unit_system))
        gamma = float(self.collision_rate.value_in_unit_system(unit.md_unit_system))
        # jnp.exp / jnp.sqrt on weak-typed Python floats evaluate in fp32
        a = np.float32(math.exp(float(np.float32(-gamma * dt))))
        e2 = np.float32(math.exp(float(np.float32(-2 * gamma * dt))))
        b = np.sqrt(np.float32(1.0) - e2, dtype=np.float32)
        return kT, dt, gamma, float(a), float(b)

    def run(self, sampler_state: SamplerState, thermodynamic_state: ThermodynamicState,
            number_of_steps: int = 5_000, nbr_list=None, progress_bar=False) -> Tuple[SamplerState, object]:
        from loguru import logger as log
        from .utils import mass_tensor, initialize_velocities

        potential = thermodynamic_state.potential
        self.box_vectors = sampler_state.box_vectors
        self.progress_bar = progress_bar
        temperature = thermodynamic_state.temperature
        log.debug("Running Langevin dynamics")
        log.debug(f"number_of_steps = {number_of_steps}")

        key = sampler_state.new_PRNG_key
        kT, dt, gamma, a, b = self._coefficients(temperature)

        need_v = (self.refresh_velocities or sampler_state._velocities is None
                  or sampler_state._velocities.shape[0] != sampler_state.positions.shape[0])
        if need_v:
            sampler_state.velocities = initialize_velocities(temperature, potential.topology, key)

        x = sampler_state.positions.clone()
        v = sampler_state.velocities.clone()
        mass = mass_tensor(potential.topology, x.device)

        fused = self.use_fused_engine and self._fused_eligible(potential, nbr_list, sampler_state)
        if nbr_list is not None:
            if fused:
                # same bookkeeping as build_from_state (integrators.py:169); the padded arrays are
                # materialised only if somebody reads them
                nbr_list._adopt_reference(sampler_state.positions, sampler_state.box_vectors)
            else:
                nbr_list.build_from_state(sampler_state)

        if fused:
            from ._engine import run_fused_langevin
            x, v, key = run_fused_langevin(self, x, v, mass, potential, nbr_list, sampler_state, kT, dt,
                                           gamma, key, int(number_of_steps))
        else:
            x, v, key = self._run_blocks(x, v, mass, potential, nbr_list, kT, dt, a, b, key,
                                         int(number_of_steps))

        log.debug("Finished running Langevin dynamics")
        updated_sampler_state = copy.deepcopy(sampler_state)
        updated_sampler_state.positions = x
        updated_sampler_state.velocities = v
        # the reference stores the loop key in a plain attribute (integrators.py:216, App. B #6)
        updated_sampler_state.current_PRNG_key = key
        return updated_sampler_state, nbr_list

    @staticmethod
    def _fused_eligible(potential, nbr_list, sampler_state) -> bool:
        from .potential import LJPotential
        from .neighbors import NeighborListNsqrd
        try:
            from . import _engine  # noqa: F401
        except ImportError:
            return False
        return (isinstance(potential, LJPotential) and isinstance(nbr_list, NeighborListNsqrd)
                and nbr_list.space.periodic and sampler_state.box_vectors is not None
                and _engine.available() and _engine.box_supported(nbr_list, sampler_state))

    # -- building-block path ---------------------------------------------------------------------------
    def _run_blocks(self, x, v, mass, potential, nbr_list, kT, dt, a, b, key, number_of_steps):
        from .neighbors import NeighborListNsqrd
        n, dev = x.shape[0], x.device
        ctx = _lib.get_context(dev)
        half_dt = float(np.float32(dt * 0.5))
        if nbr_list is not None and nbr_list.space.periodic:
            lx, ly, lz = (float(t) for t in np.asarray(self.box_vectors.detach().cpu().numpy(), dtype=np.float32).diagonal())
            wrap = 1
        else:
            lx = ly = lz = 1.0
            wrap = 0
        has_check = isinstance(nbr_list, NeighborListNsqrd)
        flag = torch.zeros((), dtype=torch.int32, device=dev) if has_check else None
        half_skin = float(np.float32(nbr_list._skin_md() / 2.0)) if has_check else 0.0

        F = potential.compute_force(x, nbr_list)
        if not isinstance(F, torch.Tensor):
            F = torch.zeros_like(x)
        F = F.contiguous()
        steps = range(number_of_steps)
        if self.progress_bar:
            from tqdm import tqdm
            steps = tqdm(steps)
        trailing = 0
        n_rebuilds = 0
        for step in steps:
            key, subkey = random.split(key)
            if has_check:
                flag.zero_()
            ctx.call("chx_baoab_update", _lib.ptr(x), _lib.ptr(v), _lib.ptr(F), _lib.ptr(mass), n, half_dt,
                     a, b, kT, int(subkey[0]), int(subkey[1]), trailing, lx, ly, lz, wrap,
                     _lib.ptr(nbr_list.ref_positions) if has_check else _lib.ptr(None), half_skin,
                     _lib.ptr(flag))
            if has_check and bool(flag.item()):
                # the list keeps a reference to the positions it was built from: hand it a snapshot
                nbr_list.build(x.clone(), self.box_vectors)
                n_rebuilds += 1
            F = potential.compute_force(x, nbr_list)
            if not isinstance(F, torch.Tensor):
                F = torch.zeros_like(x)
            F = F.contiguous()
            trailing = 1
            elapsed_step = step + self._move_iteration * number_of_steps
            if elapsed_step % self.report_interval == 0:
                if hasattr(self, "reporter") and self.reporter is not None:
                    self._report(x, potential, nbr_list, step, self._move_iteration, elapsed_step)
                if self.save_traj_in_memory:
                    self.traj.append(x.clone())
        if trailing:
            ctx.call("chx_kick", _lib.ptr(v), _lib.ptr(F), _lib.ptr(mass), n, half_dt)
        self.last_run_stats = {"path": "blocks", "rebuilds": n_rebuilds}
        return x, v, key

    def _wrap_and_rebuild_neighborlist(self, x, nbr_list):
        """`integrators.py:220-243` (kept for API parity; the run loop fuses wrap+check)."""
        x = nbr_list.space.wrap(x, self.box_vectors)
        if nbr_list.check(x):
            nbr_list.build(x, self.box_vectors)
        return x, nbr_list

    def _report(self, x, potential, nbr_list, step: int, iteration: int, elapsed_step: int, energy=None):
        d = {
            "positions": x.clone(),
            "potential_energy": potential.compute_energy(x, nbr_list) if energy is None else energy,
            "step": step,
            "iteration": iteration,
            "elapsed_step": elapsed_step,
        }
        if nbr_list is not None:
            d["box_vectors"] = nbr_list.box_vectors
        self.reporter.report(d)
