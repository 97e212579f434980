"""MultiStateSampler with chiron's interface (`chiron/multistate.py`), replica-sharded.

What the reference does per iteration (`multistate.py:563-599`): mix replicas (a no-op stub,
`:447-495`), propagate every replica serially through its MCMCSampler (`:497-510`), evaluate the
R x K matrix of reduced potentials with R*K full energy calls (`:512-531`), report, run MBAR.

Here the same methods exist with the same names and return values, and two things are new
(SURVEY.md section 8e):

* **replica sharding** -- with `torch.distributed` initialised, rank g owns a contiguous block of
  replicas; it propagates only those and contributes their rows of the energy matrix, one
  `all_gather` per sweep (NCCL on GPUs, gloo in CPU tests) gives every rank the full matrix;
* **batched propagation** -- when every replica is LJPotential + NeighborListNsqrd +
  LangevinDynamicsMove on the same system, the rank's replicas advance in ONE fused-engine launch
  per step (replica = blockIdx.y, per-replica key and temperature) and the energy matrix of a
  temperature ladder comes from one batched energy kernel (u_kl = beta_l U_k + beta_l p_l V_k).

Replica exchange proper (`exchange="neighbors"`) is new design (the reference has none): even/odd
neighbour swaps decided identically on every rank from a shared counter-based key, only state
indices move, velocities are rescaled by sqrt(T_new / T_old).  `exchange=None` (default)
reproduces the reference: `_mix_replicas` changes nothing.
"""
import copy
import os
from typing import List, Optional

import numpy as np
import torch

from . import random, unit
from .mcmc import LangevinDynamicsMove, MCMCSampler
from .neighbors import PairsBase
from .reporters import MultistateReporter
from .states import SamplerState, ThermodynamicState, calculate_reduced_potential_at_states


# ---------------------------------------------------------------------------------------------------
# host logic shared by every rank (pure NumPy: covered by the gloo world_size-2 CPU tests)
# ---------------------------------------------------------------------------------------------------
def shard_bounds(n_replicas: int, world_size: int, rank: int):
    """Contiguous block of replicas owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_replicas, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_ids(n_replicas: int, world_size: int, rank: int, policy: str = "contiguous") -> np.ndarray:
    """Replica ids owned by `rank`.  "contiguous": the block of `shard_bounds`; "strided": rank, rank + world, ...
    -- with a monotone temperature ladder every rank then holds cold and hot replicas alike, which evens out the
    table rebuilds the hot ones cause (sizes differ by at most one either way)."""
    if policy == "strided":
        return np.arange(rank, n_replicas, world_size, dtype=int)
    lo, hi = shard_bounds(n_replicas, world_size, rank)
    return np.arange(lo, hi, dtype=int)


def gather_rows(local_rows: np.ndarray, n_replicas: int, group=None, policy: str = "contiguous") -> np.ndarray:
    """All-gather the (n_local, K) float64 rows of every rank into the (R, K) matrix.
    Single process: returns the input.  Uses the backend of the default process group: NCCL moves
    the rows through device memory over NVLink, gloo through host memory."""
    import torch.distributed as dist
    local_rows = np.ascontiguousarray(local_rows, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_rows
    world = dist.get_world_size(group)
    K = local_rows.shape[1]
    ids = [shard_ids(n_replicas, world, r, policy) for r in range(world)]
    n_max = max(len(i) for i in ids)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    send = torch.zeros((n_max, K), dtype=torch.float64, device=dev)
    send[:local_rows.shape[0]] = torch.from_numpy(local_rows).to(dev)
    recv = torch.empty((world * n_max, K), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(world, n_max, K)
    out = np.empty((n_replicas, K), dtype=np.float64)
    for r in range(world):
        out[ids[r]] = recv[r, :len(ids[r])]
    return out


class _RowGatherer:
    """All-gather of the per-rank rows of the R x K reduced-potential matrix with the rows ALREADY ON THE
    DEVICE (NCCL): the batched energy kernel's output is scaled into a persistent send buffer, one
    `all_gather_into_tensor` runs on the compute stream's NCCL queue, and the only host transfer of the sweep
    is one D2H copy of the gathered (R, K) matrix into pinned memory.  With gloo (CPU tests) the same calls run
    on host tensors."""

    def __init__(self, n_replicas: int, K: int, group=None, policy: str = "contiguous"):
        import torch.distributed as dist
        self.R, self.K, self.group = n_replicas, K, group
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = dist.get_world_size(group) if self.dist else 1
        self.rank = dist.get_rank(group) if self.dist else 0
        self.ids = [shard_ids(n_replicas, self.world, r, policy) for r in range(self.world)]
        self.n_max = max(len(i) for i in self.ids)
        self.send = self.recv = self.host = None

    def _buffers(self, like: torch.Tensor):
        backend_dev = like.device
        if self.dist is not None and self.dist.get_backend(self.group) != "nccl":
            backend_dev = torch.device("cpu")
        if self.send is None or self.send.device != backend_dev:
            self.send = torch.zeros((self.n_max, self.K), dtype=torch.float64, device=backend_dev)
            self.recv = torch.empty((self.world * self.n_max, self.K), dtype=torch.float64, device=backend_dev)
            self.host = torch.empty((self.world * self.n_max, self.K), dtype=torch.float64,
                                    pin_memory=torch.cuda.is_available())
        return backend_dev

    def __call__(self, rows: torch.Tensor, after_enqueue=None) -> np.ndarray:
        """`after_enqueue`: called once the collective and the D2H copy are in the stream, before the host waits for
        them (the engines enqueue their table pre-build there, behind the collective instead of in front of it)."""
        dev = self._buffers(rows)
        n_local = rows.shape[0]
        if self.world == 1:
            self.host[:n_local].copy_(rows, non_blocking=True)
            if after_enqueue is not None:
                after_enqueue()
            if rows.is_cuda:
                torch.cuda.current_stream(rows.device).synchronize()
            return self.host[:n_local].numpy().copy()
        self.send[:n_local].copy_(rows if rows.device == dev else rows.to(dev), non_blocking=True)
        self.dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        self.host.copy_(self.recv, non_blocking=True)
        if after_enqueue is not None:
            after_enqueue()
        if self.recv.is_cuda:
            torch.cuda.current_stream(self.recv.device).synchronize()
        flat = self.host.numpy().reshape(self.world, self.n_max, self.K)
        out = np.empty((self.R, self.K), dtype=np.float64)
        for r, ids in enumerate(self.ids):
            out[ids] = flat[r, :len(ids)]
        return out


def neighbor_swaps(u_rk: np.ndarray, replica_states: np.ndarray, iteration: int, seed: int,
                   n_accepted: Optional[np.ndarray] = None, n_proposed: Optional[np.ndarray] = None) -> np.ndarray:
    """One round of neighbour replica exchange.  State pairs (s, s+1) with s = iteration mod 2, +2, ...;
    the replicas i, j sitting in them swap with probability min(1, exp(-(u[i,s+1] + u[j,s]) + u[i,s] + u[j,s+1])).
    The uniforms come from `jax.random.uniform(fold(seed, iteration), (n_pairs,))` (legacy threefry),
    so every rank reaches the same decisions without communication.  Returns the new state of
    every replica."""
    states = np.array(replica_states, dtype=int)
    K = u_rk.shape[1]
    replica_at = np.empty(K, dtype=int)
    replica_at[states] = np.arange(states.size)
    s = np.arange(int(iteration) % 2, K - 1, 2)
    if s.size == 0:
        return states
    key = random.PRNGKey((int(seed) << 32) ^ int(iteration))
    u01 = random.uniform_host_n(key, s.size)
    # the pairs of one round are disjoint, so all decisions are taken at once
    i, j = replica_at[s], replica_at[s + 1]
    log_p = -(u_rk[i, s + 1] + u_rk[j, s]) + u_rk[i, s] + u_rk[j, s + 1]
    with np.errstate(over="ignore"):
        accept = (log_p >= 0.0) | (u01 < np.exp(log_p))
    if n_proposed is not None:
        n_proposed[s, s + 1] += 1
        n_proposed[s + 1, s] += 1
    sa = s[accept]
    states[i[accept]], states[j[accept]] = sa + 1, sa
    if n_accepted is not None:
        n_accepted[sa, sa + 1] += 1
        n_accepted[sa + 1, sa] += 1
    return states


# ---------------------------------------------------------------------------------------------------
class MultiStateSampler:
    def __init__(self, mcmc_sampler: MCMCSampler, reporter: MultistateReporter, exchange: Optional[str] = None,
                 exchange_seed: int = 0, mcmc_iterations_per_sweep: Optional[int] = None,
                 sharding: str = "strided"):
        from .analysis import MBAREstimator
        self._thermodynamic_states = None
        self._unsampled_states = None
        self._sampler_states = None
        self._replica_thermodynamic_states = None
        self._iteration = None
        self._energy_thermodynamic_states = None
        self._neighborhoods = None
        self._n_accepted_matrix = None
        self._n_proposed_matrix = None
        self._nbr_lists = None
        self._reporter = reporter
        self._metadata = None
        self._mcmc_sampler = copy.deepcopy(mcmc_sampler)
        self._online_estimator = None
        self._offline_estimator = MBAREstimator()
        if exchange not in (None, "neighbors"):
            raise ValueError("exchange must be None (reference behaviour) or 'neighbors'")
        if sharding not in ("strided", "contiguous"):
            raise ValueError("sharding must be 'strided' or 'contiguous'")
        self.sharding = sharding          # which replicas a rank owns when the run is sharded (shard_ids)
        self.exchange = exchange
        self.exchange_seed = int(exchange_seed)
        # the reference hands the multistate n_iterations to every MCMCSampler.run as ITS n_iterations
        # (`multistate.py:441-443`); None replicates that, an int overrides it
        self.mcmc_iterations_per_sweep = mcmc_iterations_per_sweep
        self.use_batched_engine = True
        self._batched = None
        self._row_gatherer = None
        self._rank, self._world = 0, 1
        # host wall time per phase of `run`, summed over sweeps (the phases end with a host read, so device time
        # is included); reset with `reset_phase_timers()`
        self.phase_seconds = {"mix": 0.0, "propagate": 0.0, "energies_and_exchange": 0.0, "report_and_analysis": 0.0,
                              "sweeps": 0}

    def reset_phase_timers(self):
        for k in self.phase_seconds:
            self.phase_seconds[k] = 0 if k == "sweeps" else 0.0

    # ---- properties (`multistate.py:86-176`) ----------------------------------------------------------
    @property
    def number_of_thermodynamic_states(self) -> int:
        return 0 if self._thermodynamic_states is None else len(self._thermodynamic_states)

    @property
    def number_of_replicas(self) -> int:
        return 0 if self._sampler_states is None else len(self._sampler_states)

    @property
    def iteration(self):
        return self._iteration

    @property
    def mcmc_sampler(self):
        return copy.deepcopy(self._mcmc_sampler)

    @property
    def sampler_states(self) -> Optional[List[SamplerState]]:
        """Copies of the SamplerStates.  In a sharded run (world_size > 1) only the replicas in
        `local_replica_ids` are propagated by this rank; the positions of the others are fetched from their
        owners here (velocities and PRNG keys of non-local replicas stay as created)."""
        if self._sampler_states is None:
            return None
        self._sync_from_engine()
        out = copy.deepcopy(self._sampler_states)
        if self._world > 1:
            mine = set(int(r) for r in self.local_replica_ids)
            xyz = self._report_positions()["positions"]
            dev = self._sampler_states[0].positions.device
            for r in range(self.number_of_replicas):
                if r not in mine:
                    out[r].positions = unit.Quantity(torch.as_tensor(xyz[r], dtype=torch.float32, device=dev),
                                                     unit.nanometer)
        return out

    @property
    def is_periodic(self):
        if self._sampler_states is None:
            return None
        self._is_periodic = self._sampler_states[0].box_vectors is not None
        return self._is_periodic

    @property
    def is_completed(self):
        return self._is_completed()

    @property
    def local_replica_ids(self) -> np.ndarray:
        """Replica ids this rank propagates (all of them in a single-process run)."""
        return shard_ids(self.number_of_replicas, self._world, self._rank, self.sharding)

    @property
    def local_replica_range(self):
        """(lo, hi) of the contiguous policy; kept for callers that shard contiguously."""
        return shard_bounds(self.number_of_replicas, self._world, self._rank)

    # ---- creation (`multistate.py:203-309`) ------------------------------------------------------------
    def create(self, thermodynamic_states: List[ThermodynamicState], sampler_states: List[SamplerState],
               nbr_lists: List[PairsBase]):
        self._online_estimator = None
        if len(thermodynamic_states) != len(sampler_states):
            raise RuntimeError("Number of thermodynamic states and sampler states must be equal.")
        self._allocate_variables(thermodynamic_states, sampler_states, nbr_lists)
        self._reporter = MultistateReporter()      # the reference overwrites the user's reporter (App. B #10)

    def _allocate_variables(self, thermodynamic_states, sampler_states, nbr_lists) -> None:
        import torch.distributed as dist
        self._thermodynamic_states = copy.deepcopy(thermodynamic_states)
        self._sampler_states = copy.deepcopy(sampler_states)
        self._nbr_lists = copy.deepcopy(nbr_lists)
        assert len(self._thermodynamic_states) == len(self._sampler_states)
        assert len(self._thermodynamic_states) == len(self._nbr_lists)
        if dist.is_available() and dist.is_initialized():
            self._rank, self._world = dist.get_rank(), dist.get_world_size()
        for r in self.local_replica_ids:       # initial build of the lists of the replicas this rank owns
            self._nbr_lists[r].build(self._sampler_states[r].positions, self._sampler_states[r].box_vectors)
        K = len(thermodynamic_states)
        self._replica_thermodynamic_states = np.arange(K, dtype=int)
        self._n_accepted_matrix = np.zeros([K, K], np.int64)
        self._n_proposed_matrix = np.zeros([K, K], np.int64)
        self._energy_thermodynamic_states = np.zeros([self.number_of_replicas, K], np.float64)
        self._traj = [[] for _ in range(self.number_of_replicas)]
        if isinstance(self._mcmc_sampler, MCMCSampler):
            self._mcmc_sampler = [copy.deepcopy(self._mcmc_sampler) for _ in range(K)]
        elif len(self._mcmc_sampler) != K:
            raise RuntimeError(f"The number of MCMCMoves ({len(self._mcmc_sampler)}) and ThermodynamicStates "
                               f"({K}) must be the same.")
        self._iteration = 0
        self._batched = None
        self._temps_K = None

    # ---- minimisation (`multistate.py:311-412`; jaxopt replaced by steepest descent on the force kernel) -----
    def _minimize_replica(self, replica_id: int, tolerance=1.0 * unit.kilojoules_per_mole / unit.nanometers,
                          max_iterations: int = 1_000) -> None:
        from loguru import logger as log
        state_id = self._replica_thermodynamic_states[replica_id]
        potential = self._thermodynamic_states[state_id].potential
        sampler_state = self._sampler_states[replica_id]
        nbr_list = self._nbr_lists[replica_id]
        tol = float(tolerance.value_in_unit_system(unit.md_unit_system)) if isinstance(tolerance, unit.Quantity) else float(tolerance)
        from .minimze import minimize_energy
        e0 = float(potential.compute_energy(sampler_state.positions, nbr_list))
        log.debug(f"Replica {replica_id + 1}/{self.number_of_replicas}: initial energy {e0:8.3f} kJ/mol")
        result = minimize_energy(sampler_state.positions, potential.compute_energy, nbr_list,
                                 maxiter=max_iterations, tolerance=tol)
        x, e = result.params, result.state.get("value", float("nan"))
        self._sampler_states[replica_id].positions = unit.Quantity(x, unit.nanometer)
        if nbr_list is not None and nbr_list.check(self._sampler_states[replica_id].positions):
            nbr_list.build(self._sampler_states[replica_id].positions, self._sampler_states[replica_id].box_vectors)
        log.debug(f"Replica {replica_id + 1}/{self.number_of_replicas}: final energy {e:8.3f} kJ/mol")

    def minimize(self, tolerance=1.0 * unit.kilojoules_per_mole / unit.nanometers, max_iterations: int = 1_000) -> None:
        if self.number_of_replicas == 0:
            raise RuntimeError("Cannot minimize replicas. The simulation must be created first.")
        for replica_id in self.local_replica_ids:
            self._minimize_replica(int(replica_id), tolerance, max_iterations)
        self._batched = None

    # ---- propagation (`multistate.py:414-445, 497-510`) ---------------------------------------------------
    def _mcmc_iterations(self):
        return self.number_of_iterations if self.mcmc_iterations_per_sweep is None else self.mcmc_iterations_per_sweep

    def _propagate_replica(self, replica_id: int):
        state_id = self._replica_thermodynamic_states[replica_id]
        (self._sampler_states[replica_id], self._thermodynamic_states[state_id],
         self._nbr_lists[replica_id]) = self._mcmc_sampler[state_id].run(
            self._sampler_states[replica_id], self._thermodynamic_states[state_id], self._mcmc_iterations(),
            self._nbr_lists[replica_id])
        self._traj[replica_id].append(self._sampler_states[replica_id].positions)

    def _propagate_replicas(self) -> None:
        from loguru import logger as log
        log.debug("Propagating all replicas...")
        batched = self._batched_engine()
        if batched is not None:
            batched.propagate(self._mcmc_iterations())
            return
        for replica_id in self.local_replica_ids:
            self._propagate_replica(int(replica_id))

    # ---- energies (`multistate.py:178-201, 512-531`) -----------------------------------------------------
    def _compute_replica_reduced_potential(self, replica_id: int) -> np.ndarray:
        # the reference passes the SamplerState as nbr_list here (App. B #8); fixed
        return calculate_reduced_potential_at_states(self._sampler_states[replica_id], self._thermodynamic_states,
                                                     self._nbr_lists[replica_id])

    def _compute_energies(self) -> None:
        from loguru import logger as log
        log.debug("Computing energy matrix for all replicas...")
        ids = self.local_replica_ids
        K = self.number_of_thermodynamic_states
        batched = self._batched_engine()
        if batched is not None:
            # rows stay on the device until the gathered matrix comes back (one D2H per sweep)
            if self._row_gatherer is None or (self._row_gatherer.R, self._row_gatherer.K) != (self.number_of_replicas, K):
                self._row_gatherer = _RowGatherer(self.number_of_replicas, K, policy=self.sharding)
            self._energy_thermodynamic_states = self._row_gatherer(batched.reduced_potentials_device(),
                                                                   after_enqueue=batched.engine.prebuild_now)
            return
        rows = np.zeros((len(ids), K))
        for k, replica_id in enumerate(ids):
            rows[k, :] = self._compute_replica_reduced_potential(int(replica_id))
        self._energy_thermodynamic_states = gather_rows(rows, self.number_of_replicas, policy=self.sharding)

    # ---- mixing (`multistate.py:447-495`) ------------------------------------------------------------------
    def _perform_swap_proposals(self):
        if self.exchange is None:
            return self._replica_thermodynamic_states        # reference: placeholder, nothing moves
        old = self._replica_thermodynamic_states
        new = neighbor_swaps(self._energy_thermodynamic_states, old, self._iteration, self.exchange_seed,
                             self._n_accepted_matrix, self._n_proposed_matrix)
        self._apply_state_change(old, new)
        return new

    def _apply_state_change(self, old, new):
        """Replicas keep their coordinates; a replica that moved to another temperature gets its
        velocities rescaled by sqrt(T_new / T_old)."""
        batched = self._batched_engine()
        if getattr(self, "_temps_K", None) is None or len(self._temps_K) != len(self._thermodynamic_states):
            # plain floats once: unit arithmetic per replica and sweep costs more than the swap decisions
            self._temps_K = np.array([float(ts.temperature.value_in_unit(unit.kelvin))
                                      for ts in self._thermodynamic_states])
        scales = {}
        for r in self.local_replica_ids:
            if old[r] != new[r]:
                scales[int(r)] = float(np.sqrt(self._temps_K[new[r]] / self._temps_K[old[r]]))
        if batched is not None:
            batched.set_states(new, scales)
        else:
            for r, s in scales.items():
                st = self._sampler_states[r]
                if st._velocities is not None:
                    st.velocities = unit.Quantity(st.velocities * s, unit.nanometer / unit.picosecond)

    def _mix_replicas(self) -> np.ndarray:
        from loguru import logger as log
        log.debug("Mixing replicas...")
        self._n_accepted_matrix[:, :] = 0
        self._n_proposed_matrix[:, :] = 0
        new_replica_states = self._perform_swap_proposals()
        self._replica_thermodynamic_states = np.asarray(new_replica_states, dtype=int)
        n_swaps_proposed = self._n_proposed_matrix.sum()
        n_swaps_accepted = self._n_accepted_matrix.sum()
        frac = n_swaps_accepted / n_swaps_proposed if n_swaps_proposed > 0 else 0.0
        log.debug(f"Accepted {n_swaps_accepted}/{n_swaps_proposed} attempted swaps ({frac * 100.0:.1f}%)")
        return new_replica_states

    # ---- main loop (`multistate.py:533-599`) -----------------------------------------------------------------
    def _is_completed(self, iteration_limit: Optional[int] = None) -> bool:
        from loguru import logger as log
        if iteration_limit is not None and self._iteration >= iteration_limit:
            log.info(f"Reached iteration limit {iteration_limit} (current iteration {self._iteration})")
            return True
        return False

    def run(self, n_iterations: int = 10) -> None:
        from loguru import logger as log
        log.info("Running simulation...")
        self.number_of_iterations = n_iterations
        if self._iteration == 0:
            self._compute_energies()
            self._report_iteration()
        import time
        ph = self.phase_seconds
        while not self._is_completed(n_iterations):
            self._iteration += 1
            log.info(f"Iteration {self._iteration}/{n_iterations}")
            t0 = time.perf_counter()
            self._mix_replicas()
            t1 = time.perf_counter()
            self._propagate_replicas()          # ends with a host read of the loop keys: device time included
            t2 = time.perf_counter()
            self._compute_energies()            # ends with the D2H copy of the gathered matrix
            t3 = time.perf_counter()
            self._report_iteration()
            self._update_analysis()
            t4 = time.perf_counter()
            ph["mix"] += t1 - t0; ph["propagate"] += t2 - t1; ph["energies_and_exchange"] += t3 - t2
            ph["report_and_analysis"] += t4 - t3; ph["sweeps"] += 1
        # every rank keeps the (identical) records in memory for the estimator; one rank owns the file
        if self._rank == 0:
            self._reporter.flush_buffer()

    # ---- reporting / analysis (`multistate.py:601-742`) ------------------------------------------------------
    def _report_energy_matrix(self):
        return {"u_kn": self._energy_thermodynamic_states.T}

    def _report_positions(self):
        """(R, N, 3) positions of ALL replicas: every rank contributes the replicas it propagates
        (one all-gather when the run is sharded)."""
        self._sync_from_engine()
        ids = self.local_replica_ids
        n_atoms = self._sampler_states[0].positions.shape[0]
        xyz = np.zeros((self.number_of_replicas, n_atoms, 3))
        for replica_id in ids:
            xyz[replica_id] = self._sampler_states[replica_id].positions.detach().cpu().numpy()
        if self._world > 1:
            flat = gather_rows(xyz[ids].reshape(len(ids), n_atoms * 3), self.number_of_replicas, policy=self.sharding)
            xyz = flat.reshape(self.number_of_replicas, n_atoms, 3)
        return {"positions": xyz}

    def _report(self, property: str):
        if property == "positions":
            return self._report_positions()
        elif property == "u_kn":
            return self._report_energy_matrix()
        elif property == "state_index":
            return {"state_index": np.array(self._replica_thermodynamic_states)}
        return None

    def _report_iteration(self):
        prop = {}
        for property in self._reporter.properties_to_report:
            p = self._report(property)
            if p:
                prop.update(p)
        self._reporter.report(prop)

    def _update_analysis(self):
        if self._offline_estimator:
            N_k = [self._iteration] * self.number_of_thermodynamic_states
            u_kn = self._reporter.get_property("u_kn")
            self._offline_estimator.initialize(u_kn=u_kn, N_k=N_k)
        elif self._online_estimator:
            self._online_estimator.update()
        else:
            raise RuntimeError("No free energy estimator provided.")

    @property
    def f_k(self) -> np.ndarray:
        if self._offline_estimator:
            return self._offline_estimator.f_k
        elif self._online_estimator:
            return self._online_estimator.f_k
        raise RuntimeError("No free energy estimator found.")

    # ---- batched engine -----------------------------------------------------------------------------------------
    def _batched_engine(self):
        if self._batched is False:
            return None
        if self._batched is None:
            self._batched = _BatchedLJReplicas.try_create(self) if self.use_batched_engine else False
            if self._batched is None:
                self._batched = False
        return self._batched or None

    def _sync_from_engine(self):
        if self._batched:
            self._batched.sync_states()


def _engine_group_count(n_local: int) -> int:
    """Engines the rank's replicas are spread over.  Two engines on two streams, each driven by its own host
    thread, fill each other's gaps (kernel ramp and drain, the latency-bound launches and host round trips of a
    table rebuild).  Measured on one B200 through `_EngineGroups.run`, 100-step runs of n x 8,192 particles
    (profiles/r02_scripts/two_engines_lockstep2.sh, _lockstep4.sh): n = 8: 2.99 -> 2.97 ms, 16: 4.47 -> 4.32 (with
    the blocks of each engine on one warp, `set_gpu_share`), 32: 7.88 -> 7.20, 64: 13.76 -> 13.49.  So two engines
    from 16 replicas per GPU up.  CHX_REMD_ENGINE_GROUPS overrides."""
    env = os.environ.get("CHX_REMD_ENGINE_GROUPS")
    if env:
        return max(1, min(int(env), n_local))
    return 2 if n_local >= 16 else 1


class _EngineGroups:
    """The LJLangevinEngine surface `_BatchedLJReplicas` uses, over G engines: replica k of the set lives in engine
    k % G (strided, so every engine holds a cross-section of the temperature ladder and goes stale at the same
    rate).  Engine 0 runs on the caller's stream in the caller's thread; engine g > 0 has a private chx context,
    its own stream and a worker thread for `run` (ctypes releases the GIL; `chx_ljmd_run` synchronises with its
    stream chunk by chunk).  Everything else is issued from the caller's thread on the caller's stream: `run`
    returns only after every side stream has drained, and the side streams wait for the caller's stream first."""

    def __init__(self, n_groups, n, box, sigma, epsilon, cutoff, skin, dt, gamma, kT, n_replicas, device):
        from . import _lib
        from ._engine import LJLangevinEngine
        self.R, self.n, self.device = int(n_replicas), int(n), torch.device(device)
        self.G = int(n_groups)
        self.members = [list(range(g, self.R, self.G)) for g in range(self.G)]
        self.engines, self.streams = [], [None]
        for g, mem in enumerate(self.members):
            ctx = None
            if g > 0:
                s = torch.cuda.Stream(device=self.device)
                self.streams.append(s)
                with torch.cuda.stream(s):
                    ctx = _lib.Context(self.device.index if self.device.index is not None else torch.cuda.current_device())
            self.engines.append(LJLangevinEngine(n, box, sigma, epsilon, cutoff, skin, dt, gamma, kT,
                                                 n_replicas=len(mem), device=self.device, ctx=ctx))
            self.engines[-1].set_gpu_share(self.G)      # split blocks over warps for 1/G of the machine
            # (staggering the engines' table rebuilds with set_chunk_phase(g, G) was measured too: no gain, the
            # shortened chunks replay no graph -- profiles/r02_scripts/two_engines_lockstep.sh)
        self._pool = None

    def close(self):
        for e in self.engines:
            e.close()
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None

    def _idx(self, g):
        return torch.as_tensor(self.members[g], device=self.device)

    def set_state(self, x, v, mass, kT_per_replica):
        x, v = x.reshape(self.R, self.n, 3), v.reshape(self.R, self.n, 3)
        for g, (eng, mem) in enumerate(zip(self.engines, self.members)):
            eng.set_state(x[g::self.G].contiguous(), v[g::self.G].contiguous(), mass, [kT_per_replica[k] for k in mem])

    def get_state(self):
        x = torch.empty((self.R, self.n, 3), dtype=torch.float32, device=self.device)
        v = torch.empty_like(x)
        for g, eng in enumerate(self.engines):
            xg, vg, _, _ = eng.get_state()
            x[g::self.G] = xg.reshape(-1, self.n, 3)
            v[g::self.G] = vg.reshape(-1, self.n, 3)
        return x, v, None, None

    def run(self, nsteps, keys):
        keys = np.asarray(keys, dtype=np.uint32).reshape(self.R, 2)
        out = np.empty_like(keys)
        main = torch.cuda.current_stream(self.device)

        def side(g):
            s = self.streams[g]
            with torch.cuda.device(self.device), torch.cuda.stream(s):
                s.wait_stream(main)
                k, _ = self.engines[g].run(nsteps, keys[g::self.G])
                s.synchronize()
            return k

        if self._pool is None and self.G > 1:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=self.G - 1, thread_name_prefix="chx-engine")
        futures = [self._pool.submit(side, g) for g in range(1, self.G)]
        try:
            out[0::self.G], _ = self.engines[0].run(nsteps, keys[0::self.G])
        finally:
            done = [f.result() for f in futures]      # re-raises a worker's error
        for g, k in enumerate(done, start=1):
            out[g::self.G] = k
        return out, None

    def energy(self):
        e = torch.empty((self.R,), dtype=torch.float64, device=self.device)
        for g, eng in enumerate(self.engines):
            e[g::self.G] = eng.energy()
        return e

    def set_prebuild(self, mode=1):
        for eng in self.engines:
            eng.set_prebuild(mode)

    def prebuild_now(self):
        for eng in self.engines:
            eng.prebuild_now()

    def set_kT(self, kT_per_replica):
        for eng, mem in zip(self.engines, self.members):
            eng.set_kT([kT_per_replica[k] for k in mem])

    def scale_velocities(self, scale_per_replica):
        for eng, mem in zip(self.engines, self.members):
            sc = [scale_per_replica[k] for k in mem]
            if any(float(c) != 1.0 for c in sc):
                eng.scale_velocities(sc)

    def step_timing(self, reset=False):
        ms, steps = 0.0, 0
        for eng in self.engines:
            m, k = eng.step_timing(reset)
            ms, steps = ms + m, steps + k
        return ms, steps

    def stats(self):
        out = None
        for eng in self.engines:
            st = eng.stats()
            if out is None:
                out = dict(st)
            else:
                for key in ("table_rebuilds", "candidate_pairs", "interacting_pairs", "launches", "reference_rebuilds",
                            "trip_slots"):
                    out[key] += st[key]
        return out


class _BatchedLJReplicas:
    """The rank's replicas in one `chx_ljmd` engine (replica = blockIdx.y) -- or, for a handful of replicas per GPU,
    in two engines that run concurrently (`_EngineGroups`)."""

    @classmethod
    def try_create(cls, ms: "MultiStateSampler"):
        from . import _engine
        from .neighbors import NeighborListNsqrd
        from .potential import LJPotential
        if not _engine.available():
            return None
        ids = [int(r) for r in ms.local_replica_ids]
        if not ids:
            return None
        ts0, nl0, st0 = ms._thermodynamic_states[0], ms._nbr_lists[ids[0]], ms._sampler_states[ids[0]]
        pot0 = ts0.potential
        if not isinstance(pot0, LJPotential) or not isinstance(nl0, NeighborListNsqrd) or not nl0.space.periodic:
            return None
        if st0.box_vectors is None:
            return None
        box0 = st0.box_lengths_host()
        sig = (pot0.sigma, pot0.epsilon, pot0.cutoff, nl0._skin_md(), st0.positions.shape[0])
        for ts in ms._thermodynamic_states:
            p = ts.potential
            if not isinstance(p, LJPotential) or (p.sigma, p.epsilon, p.cutoff) != sig[:3] or ts.temperature is None:
                return None
        move = None
        for sampler in ms._mcmc_sampler:
            sched = sampler.move.move_schedule
            if len(sched) != 1 or not isinstance(sched[0][1], LangevinDynamicsMove):
                return None
            m = sched[0][1]
            if m.reporter is not None or m.save_traj_in_memory:
                return None
            par = (float(m.timestep.value_in_unit_system(unit.md_unit_system)),
                   float(m.collision_rate.value_in_unit_system(unit.md_unit_system)), int(m.number_of_moves),
                   bool(m.integrator.refresh_velocities))
            if move is None:
                move = par
            elif move != par:
                return None
        for r in ids:
            st, nl = ms._sampler_states[r], ms._nbr_lists[r]
            if (st.box_vectors is None or st.box_lengths_host() != box0 or st.positions.shape[0] != sig[4]
                    or not isinstance(nl, NeighborListNsqrd) or nl._skin_md() != sig[3] or nl._cutoff_md() != sig[2]):
                return None
        return cls(ms, ids, box0, sig, move)

    def __init__(self, ms, ids, box, sig, move):
        from ._engine import LJLangevinEngine
        from .utils import initialize_velocities, kT_md, mass_tensor
        self.ms, self.ids = ms, list(ids)      # engine replica k = replica ids[k]
        self.dt, self.gamma, self.nsteps, self.refresh = move
        self.n = sig[4]
        st0 = ms._sampler_states[self.ids[0]]
        dev = st0.positions.device
        self.kT_of_state = [kT_md(ts.temperature) for ts in ms._thermodynamic_states]
        states = ms._replica_thermodynamic_states
        kts = [self.kT_of_state[states[r]] for r in self.ids]
        groups = _engine_group_count(len(self.ids))
        if groups > 1:
            self.engine = _EngineGroups(groups, self.n, box, sig[0], sig[1], sig[2], sig[3], self.dt, self.gamma,
                                        kts[0], len(self.ids), dev)
        else:
            self.engine = LJLangevinEngine(self.n, box, sig[0], sig[1], sig[2], sig[3], self.dt, self.gamma, kts[0],
                                           n_replicas=len(self.ids), device=dev)
        # the rebuild a sweep's propagation ends on runs beside the exchange phase (CHX_REMD_PREBUILD=0: off)
        # (2: enqueued by `_compute_energies` right behind the all-gather, 1: by the run itself)
        self.engine.set_prebuild(int(os.environ.get("CHX_REMD_PREBUILD", "2")))
        topology = ms._thermodynamic_states[0].potential.topology
        mass = mass_tensor(topology, dev)
        xs, vs = [], []
        self._pending_v_keys = {}
        for r in self.ids:
            st = ms._sampler_states[r]
            xs.append(st.positions)
            if st._velocities is None or st.velocities.shape[0] != self.n:
                vs.append(None)
            else:
                vs.append(st.velocities)
        self._missing_v = [v is None for v in vs]
        vs = [v if v is not None else torch.zeros_like(xs[0]) for v in vs]
        self.engine.set_state(torch.stack(xs).contiguous(), torch.stack(vs).contiguous(), mass, kts)
        self.mass = mass
        self.topology = topology
        self.box = box
        self.volume = float(box[0] * box[1] * box[2])
        self._dirty = False
        self._beta_mol = None
        self._beta_pv = None

    def propagate(self, n_mcmc_iterations: int):
        """`MCMCSampler.run` x n_iterations of the single LangevinDynamicsMove for every local replica:
        per run the loop key is `sampler_state.new_PRNG_key` (integrators.py:124) and the state key
        advances by that one split (App. B #6)."""
        from .utils import initialize_velocities
        ms = self.ms
        for _ in range(int(n_mcmc_iterations)):
            # SamplerState.new_PRNG_key of every local replica in one call (states.py:150-154)
            cur = np.stack([np.asarray(ms._sampler_states[r]._current_PRNG_key, dtype=np.uint32)
                            for r in self.ids])
            carried, keys = random.split_many(cur)
            for k, r in enumerate(self.ids):
                ms._sampler_states[r]._current_PRNG_key = carried[k]
            if self.refresh or any(self._missing_v):
                x, v, _, _ = self.engine.get_state()
                x = x.reshape(len(self.ids), self.n, 3)
                v = v.reshape(len(self.ids), self.n, 3).clone()
                states = ms._replica_thermodynamic_states
                for k, r in enumerate(self.ids):
                    if self.refresh or self._missing_v[k]:
                        T = ms._thermodynamic_states[states[r]].temperature
                        v[k] = initialize_velocities(T, self.topology, keys[k]) \
                            .value_in_unit_system(unit.md_unit_system)
                kts = [self.kT_of_state[states[r]] for r in self.ids]
                self.engine.set_state(x.contiguous(), v.contiguous(), self.mass, kts)
                self._missing_v = [False] * len(self.ids)
            self.engine.run(self.nsteps, keys)
            self._dirty = True

    def _betas(self):
        if self._beta_mol is None:
            ms = self.ms
            one = unit.Quantity(1.0, unit.kilojoule_per_mole) / unit.AVOGADRO_CONSTANT_NA
            self._beta_mol = np.array([float(ts.beta * one) for ts in ms._thermodynamic_states])
            self._beta_pv = np.array([float(ts.beta * (ts.pressure * (self.volume * unit.nanometer ** 3)))
                                      if ts.pressure is not None else 0.0 for ts in ms._thermodynamic_states])
            dev = self.engine.device
            self._beta_mol_dev = torch.from_numpy(self._beta_mol).to(dev)
            self._beta_pv_dev = torch.from_numpy(self._beta_pv).to(dev)
        return self._beta_mol, self._beta_pv

    def reduced_potentials_device(self) -> torch.Tensor:
        """(n_local, K) float64 on the device: u_kl = beta_l (U_k + p_l V) from one batched energy kernel
        (`states.py:302-325` evaluated for all states at once); nothing touches the host."""
        self._betas()
        U = self.engine.energy()                               # kJ/mol per local replica, device
        return torch.addcmul(self._beta_pv_dev[None, :], U[:, None], self._beta_mol_dev[None, :])

    def reduced_potentials(self) -> np.ndarray:
        """Host copy of `reduced_potentials_device`."""
        return self.reduced_potentials_device().cpu().numpy()

    def set_states(self, new_states, velocity_scales: dict):
        kts = [self.kT_of_state[new_states[r]] for r in self.ids]
        self.engine.set_kT(kts)
        if velocity_scales:
            s = [velocity_scales.get(r, 1.0) for r in self.ids]
            self.engine.scale_velocities(s)
            self._dirty = True

    def sync_states(self):
        """Copy positions / velocities of the engine back into the SamplerState objects."""
        if not self._dirty:
            return
        x, v, _, _ = self.engine.get_state()
        x = x.reshape(len(self.ids), self.n, 3)
        v = v.reshape(len(self.ids), self.n, 3)
        for k, r in enumerate(self.ids):
            st = self.ms._sampler_states[r]
            st.positions = unit.Quantity(x[k].clone(), unit.nanometer)
            st.velocities = unit.Quantity(v[k].clone(), unit.nanometer / unit.picosecond)
        self._dirty = False
