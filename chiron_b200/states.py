"""SamplerState / ThermodynamicState with chiron's interface (`chiron/states.py`).

Arrays handed out by the properties are float32 CUDA tensors in md units (nm, ps, kJ/mol) instead
of `jnp` arrays; they are cached on the device so repeated access does not re-upload
(the reference converts on every access, `chiron/states.py:156-163`).
"""
from typing import List, Optional

import numpy as np
import torch

from . import _lib, random, unit


def _to_md_tensor(q: unit.Quantity) -> torch.Tensor:
    v = q.value_in_unit_system(unit.md_unit_system)
    return _lib.as_device_f32(v)


class SamplerState:
    """Positions, velocities, box vectors and the PRNG key (`chiron/states.py:8-174`)."""

    def __init__(self, positions, current_PRNG_key, velocities=None, box_vectors=None) -> None:
        if not isinstance(positions, unit.Quantity):
            raise TypeError(f"positions must be a unit.Quantity, got {type(positions)} instead.")
        if velocities is not None and not isinstance(velocities, unit.Quantity):
            raise TypeError(f"velocities must be a unit.Quantity, got {type(velocities)} instead.")
        if box_vectors is not None and not isinstance(box_vectors, unit.Quantity):
            if isinstance(box_vectors, list):
                try:
                    box_vectors = self._convert_from_openmm_box(box_vectors)
                except Exception:
                    raise TypeError(f"Unable to parse box_vectors {box_vectors}.")
            else:
                raise TypeError(
                    f"box_vectors must be a unit.Quantity or openMM box, got {type(box_vectors)} instead.")
        if not positions.unit.is_compatible(unit.nanometer):
            raise ValueError(f"positions must have units of distance, got {positions.unit} instead.")
        if velocities is not None and not velocities.unit.is_compatible(unit.nanometer / unit.picosecond):
            raise ValueError(f"velocities must have units of distance/time, got {velocities.unit} instead.")
        if box_vectors is not None and not box_vectors.unit.is_compatible(unit.nanometer):
            raise ValueError(f"box_vectors must have units of distance, got {box_vectors.unit} instead.")
        if box_vectors is not None and tuple(box_vectors.shape) != (3, 3):
            raise ValueError(f"box_vectors must be a 3x3 array, got {box_vectors.shape} instead.")
        if velocities is not None and tuple(positions.shape) != tuple(velocities.shape):
            raise ValueError(
                f"positions and velocities must have the same shape, got {positions.shape} and {velocities.shape} instead.")
        if current_PRNG_key is None:
            raise ValueError("random_seed must be set.")
        self._positions = positions
        self._velocities = velocities
        self._current_PRNG_key = current_PRNG_key
        self._box_vectors = box_vectors
        self._distance_unit = unit.nanometer
        self._time_unit = unit.picosecond
        self._cache = {}

    # -- device views ---------------------------------------------------------------------------------
    def _view(self, name):
        q = getattr(self, "_" + name)
        if q is None:
            return None
        hit = self._cache.get(name)
        if hit is not None and hit[0] is q:
            return hit[1]
        t = _to_md_tensor(q)
        self._cache[name] = (q, t)
        return t

    @property
    def number_of_particles(self) -> int:
        return self._positions.shape[0]

    @property
    def positions(self) -> torch.Tensor:
        return self._view("positions")

    @property
    def velocities(self) -> Optional[torch.Tensor]:
        return self._view("velocities")

    @property
    def box_vectors(self) -> Optional[torch.Tensor]:
        return self._view("box_vectors")

    @positions.setter
    def positions(self, x0) -> None:
        self._positions = x0 if isinstance(x0, unit.Quantity) else unit.Quantity(x0, self._distance_unit)

    @box_vectors.setter
    def box_vectors(self, box_vectors) -> None:
        self._box_vectors = box_vectors if isinstance(box_vectors, unit.Quantity) else \
            unit.Quantity(box_vectors, self._distance_unit)

    @velocities.setter
    def velocities(self, velocities) -> None:
        if tuple(velocities.shape) != tuple(self._positions.shape):
            raise ValueError(
                f"velocities must have the same shape as positions, got {velocities.shape} and {self._positions.shape} instead.")
        self._velocities = velocities if isinstance(velocities, unit.Quantity) else \
            unit.Quantity(velocities, self._distance_unit / self._time_unit)

    @property
    def distance_unit(self):
        return self._distance_unit

    def velocity_unit(self):
        return self._distance_unit / self._time_unit

    @property
    def new_PRNG_key(self):
        """Split the stored key, keep row 0, hand out row 1 (`chiron/states.py:150-154`)."""
        key, subkey = random.split(self._current_PRNG_key)
        self._current_PRNG_key = key
        return subkey

    def box_lengths_host(self):
        """(lx, ly, lz) as Python floats, or None.  One small D2H read, cached per box object."""
        if self._box_vectors is None:
            return None
        hit = self._cache.get("box_host")
        if hit is not None and hit[0] is self._box_vectors:
            return hit[1]
        b = self._box_vectors.value_in_unit_system(unit.md_unit_system)
        if isinstance(b, torch.Tensor):
            b = b.detach().cpu().numpy()
        b = np.asarray(b, dtype=np.float32)
        out = (float(b[0, 0]), float(b[1, 1]), float(b[2, 2]))
        self._cache["box_host"] = (self._box_vectors, out)
        return out

    def _convert_from_openmm_box(self, openmm_box_vectors: List) -> unit.Quantity:
        u0 = openmm_box_vectors[0].unit
        rows = [[float(openmm_box_vectors[i][j].value_in_unit(u0)) for j in range(3)] for i in range(3)]
        return unit.Quantity(np.array(rows), u0)


class ThermodynamicState:
    """Temperature / pressure / volume and the reduced potential (`chiron/states.py:177-329`)."""

    def __init__(self, potential, temperature=None, volume=None, pressure=None):
        self.potential = potential
        if temperature is not None and not isinstance(temperature, unit.Quantity):
            raise TypeError(f"temperature must be a unit.Quantity, got {type(temperature)} instead.")
        elif temperature is not None and not temperature.unit.is_compatible(unit.kelvin):
            raise ValueError(f"temperature must have units of temperature, got {temperature.unit} instead.")
        if volume is not None and not isinstance(volume, unit.Quantity):
            raise TypeError(f"volume must be a unit.Quantity, got {type(volume)} instead.")
        elif volume is not None and not volume.unit.is_compatible(unit.nanometer ** 3):
            raise ValueError(f"volume must have units of distance**3, got {volume.unit} instead.")
        if pressure is not None and not isinstance(pressure, unit.Quantity):
            raise TypeError(f"pressure must be a unit.Quantity, got {type(pressure)} instead.")
        elif pressure is not None and not pressure.unit.is_compatible(unit.atmosphere):
            raise ValueError(f"pressure must have units of pressure, got {pressure.unit} instead.")
        self.temperature = temperature
        self.beta = 1.0 / (unit.BOLTZMANN_CONSTANT_kB * self.temperature) if temperature is not None else None
        self.volume = volume
        self.pressure = pressure
        from .utils import get_nr_of_particles
        self.nr_of_particles = get_nr_of_particles(self.potential.topology)
        self._check_completeness()

    def check_variables(self):
        return [v for v in ("temperature", "volume", "pressure") if getattr(self, v) is not None]

    def _check_completeness(self):
        from loguru import logger as log
        set_variables = self.check_variables()
        if not set_variables:
            log.info("No variables are set.")
        for var in set_variables:
            log.info(f"{var} is set.")
        if self.temperature and self.volume and self.nr_of_particles:
            log.info("NVT ensemble simulated.")
        if self.temperature and self.pressure and self.nr_of_particles:
            log.info("NpT ensemble is simulated.")

    def get_reduced_potential(self, sampler_state: SamplerState, nbr_list=None):
        """u = beta (U + p V) (`chiron/states.py:275-325`).  Stays on the device as a 0-d fp32 tensor
        when the potential returns one; the op chain mirrors the reference's fp32 arithmetic."""
        if self.beta is None:
            self.beta = 1.0 / (unit.BOLTZMANN_CONSTANT_kB * (self.temperature * unit.kelvin))
        energy = self.potential.compute_energy(sampler_state.positions, nbr_list)
        reduced_potential = unit.Quantity(energy, unit.kilojoule_per_mole) / unit.AVOGADRO_CONSTANT_NA
        if self.pressure is not None:
            box = sampler_state.box_vectors
            self.volume = (box[0][0] * box[1][1] * box[2][2]) * unit.nanometer ** 3
            reduced_potential += self.pressure * self.volume
        return self.beta * reduced_potential

    def kT_to_kJ_per_mol(self, energy):
        energy = energy * unit.AVOGADRO_CONSTANT_NA
        return energy / self.beta


def calculate_reduced_potential_at_states(sampler_state: SamplerState,
                                          thermodynamic_states: List[ThermodynamicState],
                                          nbr_list=None):
    """Reduced potential of one configuration at every thermodynamic state
    (`chiron/states.py:335-366`); returns a float64 NumPy vector."""
    out = np.zeros(len(thermodynamic_states))
    for k, state in enumerate(thermodynamic_states):
        out[k] = float(state.get_reduced_potential(sampler_state, nbr_list))
    return out
