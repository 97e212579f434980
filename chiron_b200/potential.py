"""Potentials with chiron's interface (`chiron/potential.py`), evaluated by libchiron_b200.

Energies are returned as 0-d float32 CUDA tensors (no host sync); forces as (N,3) float32 CUDA
tensors.  The reference obtains forces by reverse-mode autodiff of the energy
(`potential.py:21-24`); here the analytic gradient is evaluated in the same kernel as the energy.
"""
import numpy as np
import torch

from . import _lib, unit
from .topology import Topology


def _check_topology(topology):
    if not isinstance(topology, (Topology, property)) and topology is not None:
        raise TypeError(f"Topology must be a Topology object or None, type(topology) = {type(topology)}")


class NeuralNetworkPotential:
    """Base class (`potential.py:7-63`)."""

    def __init__(self, model, **kwargs):
        from loguru import logger as log
        if model is None:
            log.warning("No model provided, using default model")
        else:
            self.model = model
            self.topology = model.potential.topology

    def compute_energy(self, positions, nbr_list=None):
        raise NotImplementedError

    def compute_force(self, positions, nbr_list=None):
        raise NotImplementedError(
            "chiron_b200 potentials provide analytic forces; subclasses must implement compute_force")

    def compute_pairlist(self, positions, cutoff):
        """Non-periodic O(N^2) pair list (`potential.py:26-63`): (distance, displacement, pairs(2,P))
        for i<j with d < cutoff, in row-major (i, j) order."""
        x = _lib.as_device_f32(positions)
        n = x.shape[0]
        iu = torch.triu_indices(n, n, offset=1, device=x.device)
        from .neighbors import OrthogonalNonPeriodicSpace
        r, d = OrthogonalNonPeriodicSpace().displacement(x[iu[0]], x[iu[1]], None)
        keep = d < float(cutoff)
        return d[keep], r[keep], iu[:, keep]


class IdealGasPotential(NeuralNetworkPotential):
    """U = 0 (`potential.py:66-127`)."""

    def __init__(self, topology: Topology):
        if not isinstance(topology, (Topology, property)) and topology is not None:
            raise TypeError(
                f"Topology must be a Topology object, a property, or None, got type(topology) = {type(topology)}")
        self.topology = topology

    def compute_energy(self, positions, nbr_list=None, debug_mode=False):
        return 0.0

    def compute_force(self, positions, nbr_list=None):
        # the reference returns the scalar 0.0 (potential.py:127); zeros of the right shape broadcast
        # identically and keep the integrator kernels shape-safe
        x = _lib.as_device_f32(positions)
        return torch.zeros_like(x)


class LJMixturePotential(NeuralNetworkPotential):
    """Lennard-Jones with PER-PARTICLE sigma / epsilon (Lorentz-Berthelot mixing), an optional energy shift at
    the cutoff and an optional switching function (OpenMM's quintic, from `switch_distance` to the cutoff) -- the
    generalisation `LJPotential`'s signature hints at (the reference takes
    one sigma and one epsilon, `potential.py:131-137`).  Same call surface as `LJPotential`; needs a built
    `NeighborListNsqrd`.  It is deliberately not a subclass of `LJPotential`: the fused Langevin engine and the
    device-resident Monte Carlo loops are single-species, so this class runs through the building blocks."""

    def __init__(self, topology: Topology, sigma: unit.Quantity, epsilon: unit.Quantity,
                 cutoff: unit.Quantity = unit.Quantity(1.0, unit.nanometer), shift: bool = False,
                 switch_distance: unit.Quantity = None):
        _check_topology(topology)
        for name, q, u in (("sigma", sigma, unit.angstrom), ("epsilon", epsilon, unit.kilocalories_per_mole),
                           ("cutoff", cutoff, unit.nanometer)):
            if not isinstance(q, unit.Quantity):
                raise TypeError(f"{name} must be a unit.Quantity, type({name}) = {type(q)}")
            if not q.unit.is_compatible(u):
                raise ValueError(f"{name} has the wrong units: {q.unit}")
        self._sigma_host = np.atleast_1d(np.asarray(sigma.value_in_unit_system(unit.md_unit_system), dtype=np.float32))
        self._epsilon_host = np.atleast_1d(np.asarray(epsilon.value_in_unit_system(unit.md_unit_system), dtype=np.float32))
        if self._sigma_host.shape != self._epsilon_host.shape or self._sigma_host.ndim != 1:
            raise ValueError("sigma and epsilon must be 1-d arrays of the same length (one entry per particle)")
        self.cutoff = cutoff.value_in_unit_system(unit.md_unit_system)
        self.shift = bool(shift)
        self.switch_distance = 0.0
        if switch_distance is not None:
            if not isinstance(switch_distance, unit.Quantity) or not switch_distance.unit.is_compatible(unit.nanometer):
                raise ValueError("switch_distance must be a unit.Quantity of length")
            self.switch_distance = float(switch_distance.value_in_unit_system(unit.md_unit_system))
            if not 0.0 < self.switch_distance < self.cutoff:
                raise ValueError("switch_distance must lie between 0 and the cutoff")
        self.topology = topology
        self._dev = {}

    def _params(self, n, dev):
        if self._sigma_host.shape[0] != n:
            raise ValueError(f"{self._sigma_host.shape[0]} per-particle parameters for {n} particles")
        hit = self._dev.get(dev)
        if hit is None:
            hit = (torch.from_numpy(self._sigma_host).to(dev), torch.from_numpy(self._epsilon_host).to(dev))
            self._dev[dev] = hit
        return hit

    def _evaluate(self, positions, nbr_list, want_energy, want_force):
        from .neighbors import NeighborListNsqrd
        x = _lib.as_device_f32(positions)
        n, dev = x.shape[0], x.device
        if not isinstance(nbr_list, NeighborListNsqrd):
            raise ValueError("LJMixturePotential needs a NeighborListNsqrd")
        if not nbr_list.is_built:
            raise ValueError("Neighborlist must be built before use")
        if nbr_list.cutoff.value_in_unit_system(unit.md_unit_system) != self.cutoff:
            raise ValueError(
                f"Neighborlist cutoff ({nbr_list.cutoff}) must be the same as the potential cutoff ({self.cutoff})")
        sig, eps = self._params(n, dev)
        energy = torch.zeros((), dtype=torch.float64, device=dev) if want_energy else None
        force = torch.empty((n, 3), dtype=torch.float32, device=dev) if want_force else None
        lx, ly, lz, periodic = nbr_list._box_args()
        _lib.get_context(dev).call("chx_lj_nlist_energy_force_mixed", _lib.ptr(x), n, lx, ly, lz, periodic,
                                   _lib.ptr(nbr_list.neighbor_list), _lib.ptr(nbr_list.n_neighbors),
                                   nbr_list.neighbor_list.shape[1], _lib.ptr(sig), _lib.ptr(eps), self.cutoff,
                                   int(self.shift), self.switch_distance, _lib.ptr(energy), _lib.ptr(force))
        return energy, force

    def compute_energy(self, positions, nbr_list=None, debug_mode=False):
        return self._evaluate(positions, nbr_list, True, False)[0].float()

    def compute_force(self, positions, nbr_list=None):
        return self._evaluate(positions, nbr_list, False, True)[1]

    def compute_energy_and_force(self, positions, nbr_list=None):
        e, f = self._evaluate(positions, nbr_list, True, True)
        return e.float(), f


class LJPotential(NeuralNetworkPotential):
    """Single-species Lennard-Jones, plain truncation at the cutoff (`potential.py:130-332`)."""

    def __init__(self, topology: Topology, sigma: unit.Quantity = 3.350 * unit.angstroms,
                 epsilon: unit.Quantity = 1.0 * unit.kilocalories_per_mole,
                 cutoff: unit.Quantity = unit.Quantity(1.0, unit.nanometer)):
        _check_topology(topology)
        if not isinstance(sigma, unit.Quantity):
            raise TypeError(f"sigma must be a unit.Quantity, type(sigma) = {type(sigma)}")
        if not isinstance(epsilon, unit.Quantity):
            raise TypeError(f"epsilon must be a unit.Quantity, type(epsilon) = {type(epsilon)}")
        if not isinstance(cutoff, unit.Quantity):
            raise TypeError(f"cutoff must be a unit.Quantity, type(cutoff) = {type(cutoff)}")
        if not sigma.unit.is_compatible(unit.angstrom):
            raise ValueError(f"sigma must have units of distance, got {sigma.unit}")
        if not epsilon.unit.is_compatible(unit.kilocalories_per_mole):
            raise ValueError(f"epsilon must have units of energy, got  {epsilon.unit}")
        if not cutoff.unit.is_compatible(unit.nanometer):
            raise ValueError(f"cutoff must have units of distance, got {cutoff.unit}")
        self.sigma = sigma.value_in_unit_system(unit.md_unit_system)
        self.epsilon = epsilon.value_in_unit_system(unit.md_unit_system)
        self.cutoff = cutoff.value_in_unit_system(unit.md_unit_system)
        self.topology = topology

    # -- dispatch ------------------------------------------------------------------------------------
    def _evaluate(self, positions, nbr_list, want_energy, want_force):
        from .neighbors import NeighborListNsqrd, PairListNsqrd
        x = _lib.as_device_f32(positions)
        n, dev = x.shape[0], x.device
        ctx = _lib.get_context(dev)
        energy = torch.zeros((), dtype=torch.float64, device=dev) if want_energy else None
        force = torch.empty((n, 3), dtype=torch.float32, device=dev) if want_force else None
        if nbr_list is None:
            # potential.py:235-258: inefficient N^2 pair list without PBC
            ctx.call("chx_lj_allpairs_energy_force", _lib.ptr(x), n, 1.0, 1.0, 1.0, 0, self.sigma,
                     self.epsilon, self.cutoff, _lib.ptr(energy), _lib.ptr(force))
            return energy, force
        if not nbr_list.is_built:
            raise ValueError("Neighborlist must be built before use")
        if isinstance(nbr_list, PairListNsqrd):
            if nbr_list.cutoff is not None and \
                    nbr_list.cutoff.value_in_unit_system(unit.md_unit_system) != self.cutoff:
                raise ValueError(
                    f"Neighborlist cutoff ({nbr_list.cutoff}) must be the same as the potential cutoff ({self.cutoff})")
            lx, ly, lz, periodic = nbr_list._box_args()
            ctx.call("chx_lj_allpairs_energy_force", _lib.ptr(x), n, lx, ly, lz, periodic, self.sigma,
                     self.epsilon, nbr_list._cutoff_md(), _lib.ptr(energy), _lib.ptr(force))
            return energy, force
        if nbr_list.cutoff.value_in_unit_system(unit.md_unit_system) != self.cutoff:
            raise ValueError(
                f"Neighborlist cutoff ({nbr_list.cutoff}) must be the same as the potential cutoff ({self.cutoff})")
        if not isinstance(nbr_list, NeighborListNsqrd):
            # duck-typed third-party list: use its calculate() like the reference does
            return self._evaluate_from_calculate(x, nbr_list, want_energy, want_force)
        lx, ly, lz, periodic = nbr_list._box_args()
        M = nbr_list.neighbor_list.shape[1]
        ctx.call("chx_lj_nlist_energy_force", _lib.ptr(x), n, lx, ly, lz, periodic,
                 _lib.ptr(nbr_list.neighbor_list), _lib.ptr(nbr_list.n_neighbors), M, self.sigma,
                 self.epsilon, self.cutoff, _lib.ptr(energy), _lib.ptr(force))
        return energy, force

    def _evaluate_from_calculate(self, x, nbr_list, want_energy, want_force):
        n_nb, pairs, mask, dist, rij = nbr_list.calculate(x)
        m = mask != 0
        q6 = (self.sigma / dist) ** 6
        energy = torch.where(m, 4 * self.epsilon * (q6 * q6 - q6), torch.zeros_like(dist)).sum().double() \
            if want_energy else None
        force = None
        if want_force:
            f = torch.where(m, 24 * (self.epsilon / (dist * dist)) * (2 * q6 * q6 - q6), torch.zeros_like(dist))
            fv = f[..., None] * rij
            force = fv.sum(dim=1)
            force.index_add_(0, pairs.reshape(-1).long(), -fv.reshape(-1, 3))
        return energy, force

    def compute_energy(self, positions, nbr_list=None, debug_mode=False):
        energy, _ = self._evaluate(positions, nbr_list, True, False)
        return energy.float()

    def compute_force(self, positions, nbr_list=None):
        _, force = self._evaluate(positions, nbr_list, False, True)
        return force

    def compute_energy_and_force(self, positions, nbr_list=None):
        """Both in one launch (the fused path the integrator uses on report steps)."""
        energy, force = self._evaluate(positions, nbr_list, True, True)
        return energy.float(), force

    def compute_force_analytical(self, positions):
        """`potential.py:302-332`: analytic force over the non-periodic pair list."""
        dist, disp, pairs = self.compute_pairlist(positions, self.cutoff)
        f = (24 * (self.epsilon / (dist * dist)) * (2 * (self.sigma / dist) ** 12 - (self.sigma / dist) ** 6))
        fv = f.reshape(-1, 1) * disp
        out = torch.zeros_like(_lib.as_device_f32(positions))
        out.index_add_(0, pairs[0], fv)
        out.index_add_(0, pairs[1], -fv)
        return out


class HarmonicOscillatorPotential(NeuralNetworkPotential):
    """U = k/2 sum (x - x0)^2 + U0 (`potential.py:335-428`)."""

    def __init__(self, topology: Topology,
                 k: unit.Quantity = 1.0 * unit.kilocalories_per_mole / unit.angstrom ** 2,
                 x0: unit.Quantity = np.array([[0.0, 0.0, 0.0]]) * unit.angstrom,
                 U0: unit.Quantity = 0.0 * unit.kilocalories_per_mole):
        _check_topology(topology)
        if not isinstance(k, unit.Quantity):
            raise TypeError(f"k must be a unit.Quantity, type(k) = {type(k)}")
        if not isinstance(x0, unit.Quantity):
            raise TypeError(f"positions must be a unit.Quantity, type(positions) = {type(x0)}")
        if not isinstance(U0, unit.Quantity):
            raise TypeError(f"U0 must be a unit.Quantity, type(U0) = {type(U0)}")
        if not k.unit.is_compatible(unit.kilocalories_per_mole / unit.angstrom ** 2):
            raise ValueError(
                f"k must be a unit.Quantity with units of energy per distance squared, k.unit = {k.unit}")
        if not x0.unit.is_compatible(unit.angstrom):
            raise ValueError(
                f"positions must be a unit.Quantity with units of distance, positions.unit = {x0.unit}")
        assert x0.shape[1] == 3, f"positions must be a NX3 vector, positions.shape = {x0.shape}"
        if not U0.unit.is_compatible(unit.kilocalories_per_mole):
            raise ValueError(f"U0 must be a unit.Quantity with units of energy, U0.unit = {U0.unit}")
        self.k = float(k.value_in_unit_system(unit.md_unit_system))
        x0_md = x0.value_in_unit_system(unit.md_unit_system)
        self._x0_host = np.asarray(x0_md.detach().cpu().numpy() if isinstance(x0_md, torch.Tensor) else x0_md,
                                   dtype=np.float32)
        self.U0 = float(U0.value_in_unit_system(unit.md_unit_system))
        self.topology = topology
        self._x0_dev = None

    @property
    def x0(self):
        if self._x0_dev is None:
            self._x0_dev = _lib.as_device_f32(self._x0_host)
        return self._x0_dev

    def _evaluate(self, positions, want_energy, want_force):
        x = _lib.as_device_f32(positions)
        if x.dim() == 1:
            x = x.view(1, 3)
        n, dev = x.shape[0], x.device
        x0 = self.x0
        if x0.shape[0] not in (1, n):
            raise ValueError(f"x0 has {x0.shape[0]} rows, positions {n}")
        energy = torch.zeros((), dtype=torch.float64, device=dev) if want_energy else None
        force = torch.empty((n, 3), dtype=torch.float32, device=dev) if want_force else None
        _lib.get_context(dev).call("chx_ho_energy_force", _lib.ptr(x), n, _lib.ptr(x0), x0.shape[0],
                                   self.k, self.U0, _lib.ptr(energy), _lib.ptr(force))
        return energy, force

    def compute_energy(self, positions, nbr_list=None):
        return self._evaluate(positions, True, False)[0].float()

    def compute_force(self, positions, nbr_list=None):
        return self._evaluate(positions, False, True)[1]

    def compute_energy_and_force(self, positions, nbr_list=None):
        e, f = self._evaluate(positions, True, True)
        return e.float(), f
