"""Seed/key handling, masses and Maxwell-Boltzmann velocities (mirrors `chiron/utils.py`)."""
import numpy as np
import torch

from . import _lib, random, unit
from .topology import Topology

kB = unit.BOLTZMANN_CONSTANT_kB * unit.AVOGADRO_CONSTANT_NA


class PRNG:
    """Process-global seed -> key splitter (`chiron/utils.py:6-38`)."""

    _key = None
    _seed = None

    @classmethod
    def set_seed(cls, seed: int) -> None:
        cls._seed = seed
        cls._key = random.PRNGKey(seed)

    @classmethod
    def get_random_key(cls):
        if cls._key is None:
            raise RuntimeError("PRNG.set_seed must be called before PRNG.get_random_key")
        key, subkey = random.split(cls._key)
        cls._key = key
        return subkey


def get_nr_of_particles(topology: Topology) -> int:
    """`chiron/utils.py:101-103`."""
    return topology.getNumAtoms()


def get_list_of_mass(topology: Topology):
    """Masses from `atom.element.mass` in amu (`chiron/utils.py:106-113`)."""
    mass = [atom.element.mass.value_in_unit(unit.amu) for atom in topology.atoms()]
    return mass * unit.amu


_MASS_CACHE = {}


def mass_tensor(topology, device=None) -> torch.Tensor:
    """Masses (amu) as a float32 CUDA tensor.  The reference walks the topology atom by atom on
    every `run` (`chiron/utils.py:106-113`); here the result is cached per topology object and
    atom count, so repeated calls cost nothing."""
    dev = torch.device(device) if device is not None else _lib.default_device()
    key = (id(topology), topology.getNumAtoms(), str(dev))
    hit = _MASS_CACHE.get(key)
    if hit is not None and hit[0]() is topology:
        return hit[1]
    m = get_list_of_mass(topology).value_in_unit_system(unit.md_unit_system)
    t = _lib.as_device_f32(np.asarray(m, dtype=np.float32), dev)
    try:
        import weakref
        if len(_MASS_CACHE) > 64:
            _MASS_CACHE.clear()
        _MASS_CACHE[key] = (weakref.ref(topology), t)
    except TypeError:
        pass
    return t


def kT_md(temperature) -> float:
    return float((kB * temperature).value_in_unit_system(unit.md_unit_system))


def initialize_velocities(temperature, topology: Topology, key):
    """v0 = sqrt(kT/m) * normal(key, (N,3)) (`chiron/utils.py:116-144`); returns a Quantity."""
    mass = mass_tensor(topology)
    n = mass.shape[0]
    ctx = _lib.get_context(mass.device)
    v = torch.empty((n, 3), dtype=torch.float32, device=mass.device)
    key = random._as_key(key)
    ctx.call("chx_init_velocities", _lib.ptr(v), _lib.ptr(mass), n, kT_md(temperature),
             int(key[0]), int(key[1]))
    return v * unit.nanometer / unit.picosecond
