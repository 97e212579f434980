"""chiron_b200 -- sm_100a implementation of chiron's particle hot path behind chiron's Python API.

    from chiron_b200 import unit
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.potential import LJPotential
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.integrators import LangevinIntegrator

All arithmetic runs in libchiron_b200.so (hand-written CUDA, C ABI in include/chiron_b200.h).
There is no CPU fallback: compute calls raise `ChironB200Error` without the library or a GPU.
"""
from . import unit  # noqa: F401
from ._lib import ChironB200Error, LIB_PATH, load_library  # noqa: F401

__version__ = "0.1.0"
