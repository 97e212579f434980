"""MCMC moves, schedule and sampler with chiron's interface (`chiron/mcmc.py`).

The control flow (two PRNG splits per Monte Carlo step, statistics, autotune and report cadence,
accept-returns-new-state / reject-returns-same-state-with-advanced-key) follows the reference;
proposals and energies run in libchiron_b200 kernels and the state copies the reference makes
with `copy.deepcopy` are replaced by tensor clones of what actually changes.
"""
import copy
import math
import os
from abc import abstractmethod
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib, random, unit
from .states import SamplerState, ThermodynamicState


class MCMCMove:
    """`mcmc.py:11-89`."""

    def __init__(self, number_of_moves: int, reporter=None, report_interval: Optional[int] = 100):
        self.number_of_moves = number_of_moves
        self.reporter = reporter
        self.report_interval = report_interval
        self._move_iteration = 0
        self._number_of_attempts_made = 0
        if self.reporter is not None:
            assert self.report_interval is not None

    @abstractmethod
    def update(self, sampler_state, thermodynamic_state, nbr_list=None):
        pass

    @property
    def number_of_attemps_made(self):
        return self._number_of_attempts_made


class LangevinDynamicsMove(MCMCMove):
    """`mcmc.py:91-199`."""

    def __init__(self, timestep=1.0 * unit.femtoseconds, collision_rate=1.0 / unit.picoseconds,
                 refresh_velocities: bool = False, reporter=None, report_interval: int = 100,
                 number_of_steps: int = 1_000, save_traj_in_memory: bool = False):
        super().__init__(number_of_moves=number_of_steps, reporter=reporter, report_interval=report_interval)
        self.timestep = timestep
        self.collision_rate = collision_rate
        self.save_traj_in_memory = save_traj_in_memory
        self.traj = []
        from .integrators import LangevinIntegrator
        self.integrator = LangevinIntegrator(
            timestep=self.timestep, collision_rate=self.collision_rate,
            refresh_velocities=refresh_velocities, report_interval=report_interval, reporter=reporter,
            save_traj_in_memory=save_traj_in_memory)

    def update(self, sampler_state, thermodynamic_state, nbr_list=None):
        assert isinstance(sampler_state, SamplerState), \
            f"Sampler state must be SamplerState, not {type(sampler_state)}"
        assert isinstance(thermodynamic_state, ThermodynamicState), \
            f"Thermodynamic state must be ThermodynamicState, not {type(thermodynamic_state)}"
        updated_sampler_state, updated_nbr_list = self.integrator.run(
            thermodynamic_state=thermodynamic_state, sampler_state=sampler_state,
            number_of_steps=self.number_of_moves, nbr_list=nbr_list)
        self._number_of_attempts_made += self.number_of_moves
        if self.save_traj_in_memory:
            self.traj.append(self.integrator.traj)
            self.integrator.traj = []
        self._move_iteration += 1
        return updated_sampler_state, thermodynamic_state, updated_nbr_list


def _shallow_state_copy(state: SamplerState) -> SamplerState:
    """New SamplerState object sharing the (immutable-by-convention) array Quantities."""
    new = copy.copy(state)
    new._cache = dict(state._cache)
    new._current_PRNG_key = np.array(state._current_PRNG_key, dtype=np.uint32)
    return new


def _copy_nbr_list(nbr_list):
    """Equivalent of the reference's deepcopy of a list before rebuilding it: build() replaces the
    arrays wholesale, so a shallow copy is enough."""
    return copy.copy(nbr_list)


class MCMove(MCMCMove):
    """Metropolis loop shared by all Monte Carlo moves (`mcmc.py:202-548`)."""

    def __init__(self, number_of_moves: int, reporter, report_interval: int = 1, autotune: bool = False,
                 autotune_interval: int = 100, acceptance_method: str = "Metropolis-Hastings") -> None:
        super().__init__(number_of_moves=number_of_moves, reporter=reporter, report_interval=report_interval)
        self.acceptance_method = acceptance_method
        self.reset_statistics()
        self.autotune = autotune
        self.autotune_interval = autotune_interval

    # Run whole `update()` calls as a device-resident loop (csrc/mc.cu) when the move supports it and
    # nothing has to be reported per step.  False keeps the reference's step-by-step control flow.
    device_loop = os.environ.get("CHX_MC_DEVICE_LOOP", "1") != "0"

    def _device_plan(self, sampler_state, thermodynamic_state, nbr_list):
        """Arguments of the device loop for this move, or None to use the step-by-step path."""
        return None

    def update(self, sampler_state, thermodynamic_state, nbr_list=None):
        self._current_reduced_potential = None
        if self.device_loop and getattr(self, "reporter", None) is None \
                and self.acceptance_method == "Metropolis-Hastings" and self.number_of_moves > 0:
            plan = self._device_plan(sampler_state, thermodynamic_state, nbr_list)
            if plan is not None:
                return self._update_device(plan, sampler_state, thermodynamic_state, nbr_list)
        for i in range(self.number_of_moves):
            sampler_state, thermodynamic_state, nbr_list = self._step(sampler_state, thermodynamic_state, nbr_list)
            self._number_of_attempts_made += 1
            if getattr(self, "reporter", None) is not None:
                if self._number_of_attempts_made % self.report_interval == 0:
                    self._report(i, self._move_iteration, self._number_of_attempts_made,
                                 self.n_accepted / self.n_proposed, sampler_state, thermodynamic_state, nbr_list)
            if self.autotune:
                if self._number_of_attempts_made % self.autotune_interval == 0 and self._number_of_attempts_made > 0:
                    self._autotune()
        self._move_iteration += 1
        return sampler_state, thermodynamic_state, nbr_list

    def _update_device(self, plan, sampler_state, thermodynamic_state, nbr_list):
        """`update()` on the device loop `chx_mc_displace_run` (include/chiron_b200.h): segments end at
        autotune boundaries; a proposal that needs a neighbour-list rebuild is made by `_step`."""
        import ctypes as C
        remaining = self.number_of_moves
        while remaining > 0:
            args, keep = plan
            x = _lib.as_device_f32(sampler_state.positions)
            dev = x.device
            ctx = _lib.get_context(dev)
            # persistent work buffers: the device loop caches its captured graph by argument addresses,
            # so the same buffers must come back on every update() (results are copied out below)
            cache = getattr(self, "_dev_bufs", None)
            if cache is None or cache[0].shape != x.shape or cache[0].device != dev:
                cache = (torch.empty_like(x), torch.empty_like(x), torch.zeros(16, dtype=torch.int32, device=dev))
                self._dev_bufs = cache
            bufs, st_dev = [cache[0], cache[1]], cache[2]
            bufs[0].copy_(x)
            st_dev.zero_()
            st = _lib.McState()
            key = sampler_state._current_PRNG_key
            st.key[0], st.key[1] = int(key[0]), int(key[1])
            st.n_accepted, st.n_proposed = int(self.n_accepted), int(self.n_proposed)
            accepted_before = int(self.n_accepted)
            halted = False
            while remaining > 0 and not halted:
                seg = remaining
                if self.autotune:
                    seg = min(seg, self.autotune_interval - self._number_of_attempts_made % self.autotune_interval)
                st.sigma_disp = float(self.displacement_sigma.value_in_unit_system(unit.md_unit_system))
                ctx.call("chx_mc_displace_run", C.byref(args), _lib.ptr(bufs[0]), _lib.ptr(bufs[1]),
                         _lib.ptr(st_dev), C.byref(st), int(seg))
                done = int(st.moves_done)
                remaining -= done
                self._number_of_attempts_made += done
                self.n_accepted, self.n_proposed = int(st.n_accepted), int(st.n_proposed)
                halted = bool(st.halt)
                if self.autotune and done > 0 and self._number_of_attempts_made % self.autotune_interval == 0:
                    self._autotune()
            new_key = np.array([st.key[0], st.key[1]], dtype=np.uint32)
            if self.n_accepted != accepted_before:
                # accept returns a new state object, reject the same one with the advanced key (mcmc.py:441-463)
                out = _shallow_state_copy(sampler_state)
                out.positions = bufs[int(st.sel)].clone()
                sampler_state = out
            sampler_state._current_PRNG_key = new_key
            if halted:
                # this proposal rebuilds the neighbour list (mcmc.py:754-759): one reference-shaped step
                self._current_reduced_potential = None
                sampler_state, thermodynamic_state, nbr_list = self._step(sampler_state, thermodynamic_state, nbr_list)
                self._number_of_attempts_made += 1
                remaining -= 1
                if self.autotune and self._number_of_attempts_made % self.autotune_interval == 0:
                    self._autotune()
                if remaining > 0:
                    plan = self._device_plan(sampler_state, thermodynamic_state, nbr_list)
                    if plan is None:
                        for _ in range(remaining):
                            sampler_state, thermodynamic_state, nbr_list = self._step(
                                sampler_state, thermodynamic_state, nbr_list)
                            self._number_of_attempts_made += 1
                        remaining = 0
        self._current_reduced_potential = None
        self._move_iteration += 1
        return sampler_state, thermodynamic_state, nbr_list

    @abstractmethod
    def _report(self, step, iteration, number_of_attempts_made, acceptance_probability, sampler_state,
                thermodynamic_state, nbr_list=None):
        pass

    @abstractmethod
    def _autotune(self):
        pass

    def _step(self, current_sampler_state, current_thermodynamic_state, current_nbr_list=None):
        """`mcmc.py:357-463`."""
        if self._current_reduced_potential is None:
            current_reduced_potential = current_thermodynamic_state.get_reduced_potential(
                current_sampler_state, current_nbr_list)
            self._current_reduced_potential = current_reduced_potential
        else:
            current_reduced_potential = self._current_reduced_potential

        (proposed_sampler_state, proposed_thermodynamic_state, proposed_reduced_potential,
         log_proposal_ratio, proposed_nbr_list) = self._propose(
            current_sampler_state, current_thermodynamic_state, current_reduced_potential, current_nbr_list)

        # one host read per Monte Carlo step: the decision is made on the host like in the reference
        log_ratio_host = float(log_proposal_ratio)
        proposed_host = float(proposed_reduced_potential)
        if math.isnan(proposed_host):
            decision = False
        else:
            decision = self._accept_or_reject(log_ratio_host, proposed_sampler_state.new_PRNG_key,
                                              acceptance_method=self.acceptance_method)
        self._update_statistics(decision)
        if decision:
            self._current_reduced_potential = proposed_reduced_potential
            return proposed_sampler_state, proposed_thermodynamic_state, proposed_nbr_list
        current_sampler_state._current_PRNG_key = proposed_sampler_state._current_PRNG_key
        return current_sampler_state, current_thermodynamic_state, current_nbr_list

    def _update_statistics(self, decision):
        if decision:
            self.n_accepted += 1
        self.n_proposed += 1

    @property
    def statistics(self):
        return dict(n_accepted=self.n_accepted, n_proposed=self.n_proposed)

    @statistics.setter
    def statistics(self, value):
        self.n_accepted = value["n_accepted"]
        self.n_proposed = value["n_proposed"]

    def reset_statistics(self):
        self.n_accepted = 0
        self.n_proposed = 0

    @abstractmethod
    def _propose(self, current_sampler_state, current_thermodynamic_state, current_reduced_potential,
                 current_nbr_list=None):
        pass

    def _accept_or_reject(self, log_proposal_ratio, key, acceptance_method):
        """`mcmc.py:531-548`: accept iff log_ratio >= 0 or uniform(key) < exp(log_ratio) (fp32)."""
        if acceptance_method == "Metropolis-Hastings":
            compare_to = random.uniform_host(key)
            lr = np.float32(log_proposal_ratio)
            with np.errstate(over="ignore"):
                return bool(-lr <= 0.0 or compare_to < np.exp(lr, dtype=np.float32))


class MonteCarloDisplacementMove(MCMove):
    """Gaussian displacement of all particles (or `atom_subset`) (`mcmc.py:551-787`).

    With `atom_subset` set, an `LJPotential` and a periodic list, the proposed energy is obtained
    from the single-pass delta-energy kernel (only the moved particles' rows are evaluated);
    `use_delta_energy=False` forces the reference's full re-evaluation.
    """

    def __init__(self, displacement_sigma=1.0 * unit.nanometer, number_of_moves: int = 100,
                 atom_subset: Optional[List[int]] = None, report_interval: int = 1, reporter=None,
                 autotune: bool = False, autotune_interval: int = 100,
                 acceptance_method="Metropolis-Hastings", use_delta_energy: bool = True):
        super().__init__(number_of_moves=number_of_moves, reporter=reporter, report_interval=report_interval,
                         autotune=autotune, autotune_interval=autotune_interval,
                         acceptance_method=acceptance_method)
        self.displacement_sigma = displacement_sigma
        self.atom_subset = atom_subset
        self.atom_subset_mask = None
        self.use_delta_energy = use_delta_energy
        self._subset_ids = None

    def _report(self, step, iteration, number_of_attempts_made, acceptance_probability, sampler_state,
                thermodynamic_state, nbr_list=None):
        potential = thermodynamic_state.potential.compute_energy(sampler_state.positions, nbr_list)
        self.reporter.report({
            "step": step,
            "iteration": iteration,
            "number_of_attempts_made": number_of_attempts_made,
            "potential_energy": potential,
            "displacement_sigma": self.displacement_sigma.value_in_unit_system(unit.md_unit_system),
            "acceptance_probability": acceptance_probability,
        })

    def _autotune(self):
        acceptance_ratio = self.n_accepted / self.n_proposed
        if acceptance_ratio > 0.6:
            self.displacement_sigma *= 1.1
        elif acceptance_ratio < 0.4:
            self.displacement_sigma /= 1.1

    def _ensure_subset(self, n, dev):
        if self.atom_subset is not None and self.atom_subset_mask is None:
            m = torch.zeros(n, dtype=torch.float32, device=dev)
            ids = torch.as_tensor(list(self.atom_subset), dtype=torch.long, device=dev)
            m[ids] = 1.0
            self.atom_subset_mask = m
            self._subset_ids = ids.to(torch.int32).contiguous()

    def _device_plan(self, sampler_state, thermodynamic_state, nbr_list):
        """chx_mc_displace_args for (potential, list) combinations the device loop covers."""
        from .neighbors import NeighborListNsqrd, PairListNsqrd
        from .potential import HarmonicOscillatorPotential, IdealGasPotential, LJPotential
        from .utils import kT_md
        ts = thermodynamic_state
        if ts.temperature is None or not torch.cuda.is_available():
            return None
        x = sampler_state.positions
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and x.shape[1] == 3):
            return None
        pot = ts.potential
        n, dev = x.shape[0], x.device
        a = _lib.McDisplaceArgs()
        keep = []
        a.n = n
        a.lx = a.ly = a.lz = 1.0
        if nbr_list is not None:
            if not isinstance(nbr_list, (NeighborListNsqrd, PairListNsqrd)) or not nbr_list.is_built:
                return None
            if nbr_list.ref_positions.shape[0] != n:
                return None
            a.periodic = int(nbr_list.space.periodic)
            if a.periodic:
                if sampler_state.box_vectors is None:
                    return None
                a.lx, a.ly, a.lz = sampler_state.box_lengths_host()
                if tuple(np.float32(v) for v in nbr_list._box_args()[:3]) != \
                        tuple(np.float32(v) for v in (a.lx, a.ly, a.lz)):
                    return None
        self._ensure_subset(n, dev)
        if self.atom_subset is not None:
            a.subset_mask = _lib.ptr(self.atom_subset_mask)
            keep.append(self.atom_subset_mask)
        if isinstance(nbr_list, NeighborListNsqrd):
            ref = _lib.as_device_f32(nbr_list.ref_positions)
            a.ref_positions, a.skin = _lib.ptr(ref), nbr_list._skin_md()
            keep.append(ref)
        if type(pot) is LJPotential:
            a.sigma, a.epsilon, a.cutoff = pot.sigma, pot.epsilon, pot.cutoff
            if isinstance(nbr_list, NeighborListNsqrd):
                if nbr_list._cutoff_md() != pot.cutoff:
                    return None   # the step-by-step path raises the reference's ValueError
                if self.use_delta_energy and self.atom_subset is not None and a.periodic \
                        and len(self.atom_subset) * 8 <= n:
                    a.potential = _lib.MC_LJ_SUBSET_DELTA
                    a.subset_ids, a.n_subset = _lib.ptr(self._subset_ids), int(self._subset_ids.shape[0])
                    keep.append(self._subset_ids)
                else:
                    a.potential = _lib.MC_LJ_NLIST
                    nl, nn = nbr_list.neighbor_list, nbr_list.n_neighbors
                    a.neighbor_list, a.n_neighbors, a.M = _lib.ptr(nl), _lib.ptr(nn), int(nl.shape[1])
                    keep += [nl, nn]
            elif isinstance(nbr_list, PairListNsqrd):
                if nbr_list.cutoff is not None and nbr_list._cutoff_md() != pot.cutoff:
                    return None
                a.potential = _lib.MC_LJ_ALLPAIRS
                a.cutoff = nbr_list._cutoff_md()
            else:
                a.potential = _lib.MC_LJ_ALLPAIRS   # potential.py:235-258: non-periodic N^2 pairs, d < cutoff
            if a.potential == _lib.MC_LJ_ALLPAIRS and n > 8192:
                return None
        elif type(pot) is HarmonicOscillatorPotential:
            x0 = pot.x0
            if x0.shape[0] not in (1, n):
                return None
            a.potential, a.x0, a.n0, a.k, a.U0 = _lib.MC_HO, _lib.ptr(x0), int(x0.shape[0]), pot.k, pot.U0
            keep.append(x0)
        elif type(pot) is IdealGasPotential:
            a.potential = _lib.MC_IDEAL
        else:
            return None
        a.beta = 1.0 / kT_md(ts.temperature)
        a.pv = 0.0
        if ts.pressure is not None:
            if sampler_state.box_vectors is None:
                return None
            lx, ly, lz = sampler_state.box_lengths_host()
            volume = (lx * ly * lz) * unit.nanometer ** 3
            a.pv = float((ts.pressure * volume * unit.AVOGADRO_CONSTANT_NA).value_in_unit_system(unit.md_unit_system))
        return a, keep

    def _propose(self, current_sampler_state, current_thermodynamic_state, current_reduced_potential,
                 current_nbr_list=None):
        x = current_sampler_state.positions
        n, dev = x.shape[0], x.device
        if self.atom_subset is not None and self.atom_subset_mask is None:
            m = torch.zeros(n, dtype=torch.float32, device=dev)
            ids = torch.as_tensor(list(self.atom_subset), dtype=torch.long, device=dev)
            m[ids] = 1.0
            self.atom_subset_mask = m
            self._subset_ids = ids.to(torch.int32).contiguous()

        key = current_sampler_state.new_PRNG_key
        sigma = float(self.displacement_sigma.value_in_unit_system(unit.md_unit_system))
        proposed_sampler_state = _shallow_state_copy(current_sampler_state)
        wrap = current_nbr_list is not None and current_nbr_list.space.periodic
        if wrap:
            lx, ly, lz = current_sampler_state.box_lengths_host()
        else:
            lx = ly = lz = 1.0
        xp = torch.empty_like(x)
        _lib.get_context(dev).call(
            "chx_mc_displace", _lib.ptr(x), n, int(key[0]), int(key[1]), sigma,
            _lib.ptr(self.atom_subset_mask), lx, ly, lz, int(wrap), _lib.ptr(xp))
        proposed_sampler_state.positions = xp

        if current_nbr_list is not None:
            if current_nbr_list.check(xp):
                proposed_nbr_list = _copy_nbr_list(current_nbr_list)
                proposed_nbr_list.build(xp, proposed_sampler_state.box_vectors)
            else:
                proposed_nbr_list = current_nbr_list
        else:
            proposed_nbr_list = None

        delta = self._delta_reduced_potential(x, xp, current_sampler_state, current_thermodynamic_state,
                                              current_nbr_list)
        if delta is not None:
            # the acceptance uses -delta itself (as the device loop does), not the rounded difference of
            # two totals
            proposed_reduced_potential = current_reduced_potential + delta
            log_proposal_ratio = -delta
        else:
            proposed_reduced_potential = current_thermodynamic_state.get_reduced_potential(
                proposed_sampler_state, proposed_nbr_list)
            log_proposal_ratio = -proposed_reduced_potential + current_reduced_potential
        return (proposed_sampler_state, current_thermodynamic_state, proposed_reduced_potential,
                log_proposal_ratio, proposed_nbr_list)

    def _delta_reduced_potential(self, x, xp, sampler_state, thermodynamic_state, nbr_list):
        """beta * (U(new) - U(old)) from the subset delta-energy kernel, or None if not applicable."""
        from .neighbors import NeighborListNsqrd
        from .potential import LJPotential
        pot = thermodynamic_state.potential
        # only for lists whose energy is truncated at pot.cutoff (a PairListNsqrd(cutoff=None) is not)
        if not (self.use_delta_energy and self.atom_subset is not None and isinstance(pot, LJPotential)
                and isinstance(nbr_list, NeighborListNsqrd) and nbr_list.space.periodic
                and thermodynamic_state.temperature is not None):
            return None
        n, dev = x.shape[0], x.device
        if len(self.atom_subset) * 8 > n:
            return None
        lx, ly, lz = sampler_state.box_lengths_host()
        delta = torch.zeros((), dtype=torch.float64, device=dev)
        _lib.get_context(dev).call(
            "chx_lj_subset_delta_energy", _lib.ptr(x), _lib.ptr(xp), n, _lib.ptr(self._subset_ids),
            int(self._subset_ids.shape[0]), lx, ly, lz, 1, pot.sigma, pot.epsilon, pot.cutoff, _lib.ptr(delta))
        from .utils import kT_md
        return (delta / kT_md(thermodynamic_state.temperature)).float()


class MonteCarloBarostatMove(MCMove):
    """Isotropic volume move (`mcmc.py:790-1009`)."""

    def __init__(self, volume_max_scale=0.01, number_of_moves: int = 100, report_interval: int = 1,
                 reporter=None, autotune: bool = False, autotune_interval: int = 100,
                 acceptance_method="Metropolis-Hastings"):
        super().__init__(number_of_moves=number_of_moves, reporter=reporter, report_interval=report_interval,
                         autotune=autotune, autotune_interval=autotune_interval,
                         acceptance_method=acceptance_method)
        self.volume_max_scale = volume_max_scale

    def _report(self, step, iteration, number_of_attempts_made, acceptance_probability, sampler_state,
                thermodynamic_state, nbr_list=None):
        potential = thermodynamic_state.potential.compute_energy(sampler_state.positions, nbr_list)
        box = sampler_state.box_vectors
        volume = box[0][0] * box[1][1] * box[2][2]
        self.reporter.report({
            "step": step,
            "iteration": iteration,
            "number_of_attempts_made": number_of_attempts_made,
            "potential_energy": potential,
            "volume": volume,
            "box_vectors": box,
            "max_volume_scale": self.volume_max_scale,
            "acceptance_probability": acceptance_probability,
        })

    def _autotune(self):
        acceptance_ratio = self.n_accepted / self.n_proposed
        if acceptance_ratio < 0.25:
            self.volume_max_scale /= 1.1
        elif acceptance_ratio > 0.75:
            self.volume_max_scale = min(self.volume_max_scale * 1.1, 0.3)

    def _device_plan(self, sampler_state, thermodynamic_state, nbr_list):
        """The barostat loop `chx_mc_barostat_run` covers LJPotential over a periodic NeighborListNsqrd."""
        from .neighbors import NeighborListNsqrd
        from .potential import LJPotential
        ts = thermodynamic_state
        if ts.temperature is None or ts.pressure is None or not torch.cuda.is_available():
            return None
        x = sampler_state.positions
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and x.shape[1] == 3):
            return None
        if type(ts.potential) is not LJPotential or not isinstance(nbr_list, NeighborListNsqrd):
            return None
        if not (nbr_list.is_built and nbr_list.space.periodic) or sampler_state.box_vectors is None:
            return None
        if nbr_list._cutoff_md() != ts.potential.cutoff or nbr_list.ref_positions.shape[0] != x.shape[0]:
            return None
        box = sampler_state.box_lengths_host()
        if tuple(np.float32(v) for v in nbr_list._box_args()[:3]) != tuple(np.float32(v) for v in box):
            return None
        if min(box) < 3.0 * nbr_list._cutoff_plus_skin_md() * (1.0 + 1e-5) or x.shape[0] > 400_000:
            return None
        return True

    def _update_device(self, plan, sampler_state, thermodynamic_state, nbr_list):
        """`update()` on `chx_mc_barostat_run` (include/chiron_b200.h): positions, box and the two list
        sets stay on the device; a move whose list would outgrow `n_max_neighbors` (or whose box gets too
        small for the cell list) is made by `_step`."""
        import ctypes as C
        from .utils import kT_md
        ts = thermodynamic_state
        pot = ts.potential
        remaining = self.number_of_moves
        while remaining > 0:
            x = _lib.as_device_f32(sampler_state.positions)
            n, dev = x.shape[0], x.device
            ctx = _lib.get_context(dev)
            # set 0 = a private copy of the current list (the loop overwrites whichever set is not current)
            sets = [tuple(t.clone() for t in (nbr_list.neighbor_list, nbr_list.neighbor_mask, nbr_list.n_neighbors))]
            sets.append(tuple(t.clone() for t in sets[0]))     # both sets start valid: rows are updated incrementally
            a = _lib.McBarostatArgs()
            a.n, a.M = n, int(sets[0][0].shape[1])
            a.sigma, a.epsilon, a.cutoff = pot.sigma, pot.epsilon, pot.cutoff
            a.cutoff_plus_skin = nbr_list._cutoff_plus_skin_md()
            for k in range(2):
                a.neighbor_list[k] = sets[k][0].data_ptr()
                a.neighbor_mask[k] = sets[k][1].data_ptr()
                a.n_neighbors[k] = sets[k][2].data_ptr()
            a.beta = 1.0 / kT_md(ts.temperature)
            a.pressure = float((ts.pressure * (1.0 * unit.nanometer ** 3) * unit.AVOGADRO_CONSTANT_NA)
                               .value_in_unit_system(unit.md_unit_system))
            a.ncell_capacity = 0
            superset = (torch.empty_like(sets[0][0]), torch.empty_like(sets[0][2]))
            a.superset_list, a.superset_nn = superset[0].data_ptr(), superset[1].data_ptr()
            bufs = [x.clone(), torch.empty_like(x)]
            st_dev = torch.zeros(16, dtype=torch.int32, device=dev)
            st = _lib.McBaroState()
            key = sampler_state._current_PRNG_key
            st.key[0], st.key[1] = int(key[0]), int(key[1])
            st.n_accepted, st.n_proposed = int(self.n_accepted), int(self.n_proposed)
            st.box[0], st.box[1], st.box[2] = sampler_state.box_lengths_host()
            accepted_before = int(self.n_accepted)
            halted = False
            moved = 0
            while remaining > 0 and not halted:
                seg = remaining
                if self.autotune:
                    seg = min(seg, self.autotune_interval - self._number_of_attempts_made % self.autotune_interval)
                st.volume_max_scale = float(self.volume_max_scale)
                ctx.call("chx_mc_barostat_run", C.byref(a), _lib.ptr(bufs[0]), _lib.ptr(bufs[1]),
                         _lib.ptr(st_dev), C.byref(st), int(seg))
                done = int(st.moves_done)
                moved += done
                remaining -= done
                self._number_of_attempts_made += done
                self.n_accepted, self.n_proposed = int(st.n_accepted), int(st.n_proposed)
                halted = bool(st.halt)
                if self.autotune and done > 0 and self._number_of_attempts_made % self.autotune_interval == 0:
                    self._autotune()
            if self.n_accepted != accepted_before:
                sel = int(st.sel)
                out = _shallow_state_copy(sampler_state)
                out.positions = bufs[sel]
                old_box = sampler_state.box_vectors
                new_box = torch.zeros((3, 3), dtype=old_box.dtype, device=old_box.device)
                for k in range(3):
                    new_box[k, k] = float(st.box[k])
                out.box_vectors = new_box
                sampler_state = out
                new_list = _copy_nbr_list(nbr_list)
                new_list.ref_positions = bufs[sel]
                new_list.box_vectors = new_box
                new_list._arrays = sets[sel]
                new_list.n_builds = nbr_list.n_builds + (self.n_accepted - accepted_before)
                nbr_list = new_list
            sampler_state._current_PRNG_key = np.array([st.key[0], st.key[1]], dtype=np.uint32)
            if moved > 0:
                ts.volume = float(st.last_volume) * unit.nanometer ** 3    # states.py:313-316, last evaluation
            if halted:
                self._current_reduced_potential = None
                sampler_state, thermodynamic_state, nbr_list = self._step(sampler_state, thermodynamic_state, nbr_list)
                self._number_of_attempts_made += 1
                remaining -= 1
                if self.autotune and self._number_of_attempts_made % self.autotune_interval == 0:
                    self._autotune()
                if remaining > 0 and self._device_plan(sampler_state, thermodynamic_state, nbr_list) is None:
                    for _ in range(remaining):
                        sampler_state, thermodynamic_state, nbr_list = self._step(
                            sampler_state, thermodynamic_state, nbr_list)
                        self._number_of_attempts_made += 1
                    remaining = 0
        self._current_reduced_potential = None
        self._move_iteration += 1
        return sampler_state, thermodynamic_state, nbr_list

    def _propose(self, current_sampler_state, current_thermodynamic_state, current_reduced_potential,
                 current_nbr_list=None):
        key = current_sampler_state.new_PRNG_key
        nr_of_atoms = current_sampler_state.number_of_particles
        x = current_sampler_state.positions
        dev = x.device
        f32 = np.float32
        lx, ly, lz = (f32(t) for t in current_sampler_state.box_lengths_host())
        # scalar arithmetic in fp32 on the host, mirroring mcmc.py:956-974
        initial_volume = f32(f32(lx * ly) * lz)
        delta_volume_max = f32(f32(self.volume_max_scale) * initial_volume)
        delta_volume = f32(random.uniform_host(key, minval=-1, maxval=1) * delta_volume_max)
        proposed_volume = f32(initial_volume + delta_volume)
        length_scaling_factor = np.power(f32(proposed_volume / initial_volume), f32(1.0 / 3.0), dtype=f32)

        proposed_sampler_state = _shallow_state_copy(current_sampler_state)
        ctx = _lib.get_context(dev)
        xp = torch.empty_like(x)
        ctx.call("chx_scale", _lib.ptr(x), x.numel(), float(length_scaling_factor), _lib.ptr(xp))
        box = current_sampler_state.box_vectors
        boxp = torch.empty_like(box)
        ctx.call("chx_scale", _lib.ptr(box.contiguous()), box.numel(), float(length_scaling_factor), _lib.ptr(boxp))
        proposed_sampler_state.positions = xp
        proposed_sampler_state.box_vectors = boxp

        proposed_nbr_list = None
        if current_nbr_list is not None:
            proposed_nbr_list = _copy_nbr_list(current_nbr_list)
            proposed_nbr_list.build(xp, boxp)

        proposed_reduced_potential = current_thermodynamic_state.get_reduced_potential(
            proposed_sampler_state, proposed_nbr_list)
        log_volume = f32(f32(nr_of_atoms) * np.log(f32(proposed_volume / initial_volume), dtype=f32))
        log_proposal_ratio = -(proposed_reduced_potential - current_reduced_potential) + float(log_volume)
        return (proposed_sampler_state, current_thermodynamic_state, proposed_reduced_potential,
                log_proposal_ratio, proposed_nbr_list)


class RotamerMove(MCMove):
    def _propose(self):
        pass


class ProtonationStateMove(MCMove):
    def _propose(self):
        pass


class TautomericStateMove(MCMove):
    def _propose(self):
        pass


# the names BASELINE.json uses (they survive in the reference's test function names only)
MetropolisDisplacementMove = MonteCarloDisplacementMove
MCBarostatMove = MonteCarloBarostatMove


class MoveSchedule:
    """`mcmc.py:1036-1071`."""

    def __init__(self, move_schedule: List[Tuple[str, MCMCMove]]) -> None:
        self.move_schedule = move_schedule
        self._validate_sequence()

    def _validate_sequence(self):
        for move_name, move_class in self.move_schedule:
            if not isinstance(move_class, MCMCMove):
                raise ValueError(f"Move {move_name} in the sequence is not available.")


class MCMCSampler:
    """`mcmc.py:1074-1155`."""

    def __init__(self, move_set: MoveSchedule):
        from loguru import logger as log
        log.info("Initializing MCMC sampler")
        self.move = move_set

    def run(self, sampler_state, thermodynamic_state, n_iterations: int = 1, nbr_list=None):
        from loguru import logger as log
        sampler_state = copy.deepcopy(sampler_state)
        thermodynamic_state = copy.deepcopy(thermodynamic_state)
        nbr_list = copy.deepcopy(nbr_list)
        log.info("Running MCMC sampler")
        for iteration in range(n_iterations):
            log.info(f"Iteration {iteration + 1}/{n_iterations}")
            for move_name, move in self.move.move_schedule:
                log.debug(f"Performing: {move_name}")
                sampler_state, thermodynamic_state, nbr_list = move.update(
                    sampler_state, thermodynamic_state, nbr_list)
        log.info("Finished running MCMC sampler")
        for _, move in self.move.move_schedule:
            if move.reporter is not None:
                move.reporter.flush_buffer()
        return sampler_state, thermodynamic_state, nbr_list
