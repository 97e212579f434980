"""Deterministic synthetic systems standing in for the `openmmtools.testsystems` classes chiron's
examples and tests use (`Examples/LJ_langevin.py:6-20`, `chiron/tests/conftest.py:16-56`).

openmmtools is not available offline and its initial coordinates are not pinned by any reference
test, so the LJ fluid is a jittered simple-cubic lattice; parameters follow
openmmtools (argon: sigma 3.4 A, epsilon 0.238 kcal/mol, mass 39.948; HarmonicOscillator:
K = 100 kcal/mol/A^2, one argon-mass particle at the origin).
"""
import numpy as np

from . import unit
from .topology import Element, Topology


def _topology(n, symbol="Ar", mass=None):
    top = Topology()
    element = Element.getBySymbol(symbol) if mass is None else Element.custom(symbol, mass)
    chain = top.addChain()
    residue = top.addResidue("system", chain)
    for _ in range(n):
        top.addAtom(symbol, element, residue)
    return top


def lattice_positions(cells, box_nm, jitter_nm=0.02, seed=1):
    """Simple-cubic lattice with `cells` = (nx, ny, nz) sites in a box (lx, ly, lz), Gaussian jitter,
    wrapped into the box.  float64 -> float32, row-major site order (x slowest)."""
    cells = np.asarray(cells, dtype=int)
    box = np.asarray(box_nm, dtype=np.float64)
    a = box / cells
    g = np.stack(np.meshgrid(*[np.arange(c) for c in cells], indexing="ij"), axis=-1).reshape(-1, 3)
    x = (g + 0.5) * a
    if jitter_nm:
        x = x + np.random.default_rng(seed).normal(0.0, jitter_nm, size=x.shape)
    x = x - np.floor(x / box) * box
    return x.astype(np.float32)


class LennardJonesFluid:
    """N argon-like particles at reduced density rho* = N sigma^3 / V on a jittered lattice."""

    def __init__(self, nparticles=1000, reduced_density=0.05, sigma=3.4 * unit.angstrom,
                 epsilon=0.238 * unit.kilocalories_per_mole, mass=39.948, cutoff=None,
                 cells=None, jitter_nm=0.02, seed=1, symbol="Ar"):
        self.sigma, self.epsilon = sigma, epsilon
        s = float(sigma.value_in_unit(unit.nanometer))
        if cells is None:
            m = int(round(nparticles ** (1.0 / 3.0)))
            if m ** 3 != nparticles:
                raise ValueError("nparticles must be a cube unless `cells` is given")
            cells = (m, m, m)
        n = int(np.prod(cells))
        volume = n * s ** 3 / reduced_density
        scale = (volume / np.prod(np.asarray(cells, dtype=np.float64))) ** (1.0 / 3.0)
        box = np.asarray(cells, dtype=np.float64) * scale
        self.n_particles = n
        self.box_lengths = box
        self.cutoff = cutoff if cutoff is not None else 3.0 * sigma
        self.positions = lattice_positions(cells, box, jitter_nm, seed) * unit.nanometer
        self.box_vectors = np.diag(box).astype(np.float32) * unit.nanometer
        self.topology = _topology(n, symbol, None if symbol == "Ar" and mass == 39.948 else mass)
        self.mass = mass


class HarmonicOscillator:
    """openmmtools.testsystems.HarmonicOscillator: one particle, K = 100 kcal/mol/A^2, at the origin."""

    def __init__(self, K=100.0 * unit.kilocalories_per_mole / unit.angstrom ** 2, mass=39.948):
        self.K = K
        self.U0 = 0.0 * unit.kilojoules_per_mole
        self.positions = np.zeros((1, 3), dtype=np.float32) * unit.nanometer
        self.topology = _topology(1)


class HarmonicOscillatorArray:
    """N independent oscillators spaced d apart along x (openmmtools' HarmonicOscillatorArray)."""

    def __init__(self, K=90.0 * unit.kilocalories_per_mole / unit.angstrom ** 2, d=1.0 * unit.nanometer,
                 N=5):
        self.K, self.N = K, N
        x = np.zeros((N, 3), dtype=np.float32)
        x[:, 0] = np.arange(N) * float(d.value_in_unit(unit.nanometer))
        self.positions = x * unit.nanometer
        self.topology = _topology(N)


class IdealGas:
    """N non-interacting particles in the ideal-gas volume at (T, P)."""

    def __init__(self, nparticles=216, temperature=298.0 * unit.kelvin, pressure=1.0 * unit.atmosphere,
                 seed=2):
        kT = unit.BOLTZMANN_CONSTANT_kB * temperature
        volume = (nparticles * kT / pressure).value_in_unit(unit.nanometer ** 3)
        L = float(volume) ** (1.0 / 3.0)
        self.positions = (np.random.default_rng(seed).random((nparticles, 3)) * L).astype(np.float32) * unit.nanometer
        self.box_vectors = (np.eye(3) * L).astype(np.float32) * unit.nanometer
        self.topology = _topology(nparticles)
        self.volume = volume * unit.nanometer ** 3
