"""Minimal stand-in for `openmm.app.Topology` / `Element`.

chiron only uses the topology to count atoms (`chiron/utils.py:101-103`) and to read
`atom.element.mass` (`chiron/utils.py:106-113`); tests build one with
`Topology.addChain/addResidue/addAtom` and `Element.getBySymbol` (`chiron/tests/test_mcmc.py:386-391`).
The real OpenMM classes are used instead when OpenMM is installed.
"""
try:  # pragma: no cover - not available in the build image
    from openmm.app import Topology, Element  # noqa: F401
except Exception:  # noqa: BLE001
    from . import unit

    _MASSES = {  # amu; Ar = 39.948 is pinned by the golden trace chiron/tests/test_mcmc.py:81-84
        "H": 1.007947, "He": 4.003, "C": 12.01078, "N": 14.00672, "O": 15.99943, "Ne": 20.1797,
        "Na": 22.98977, "Cl": 35.4532, "Ar": 39.948, "Kr": 83.798, "Xe": 131.293,
    }

    class Element:
        _by_symbol = {}

        def __init__(self, symbol, mass):
            self.symbol = symbol
            self.name = symbol
            self.mass = mass * unit.dalton

        @classmethod
        def getBySymbol(cls, symbol):
            if symbol not in cls._by_symbol:
                cls._by_symbol[symbol] = Element(symbol, _MASSES[symbol])
            return cls._by_symbol[symbol]

        @classmethod
        def custom(cls, symbol, mass_amu):
            return Element(symbol, float(mass_amu))

    class _Atom:
        def __init__(self, name, element, index, residue):
            self.name, self.element, self.index, self.residue = name, element, index, residue

    class _Residue:
        def __init__(self, name, chain):
            self.name, self.chain = name, chain

    class _Chain:
        pass

    class Topology:
        def __init__(self):
            self._atoms = []

        def addChain(self, id=None):
            return _Chain()

        def addResidue(self, name, chain, id=None):
            return _Residue(name, chain)

        def addAtom(self, name, element, residue, id=None):
            a = _Atom(name, element, len(self._atoms), residue)
            self._atoms.append(a)
            return a

        def __deepcopy__(self, memo):
            # atoms are immutable records: MCMCSampler.run deep-copies its inputs (mcmc.py:1134-1136) and
            # must not spend 100 ms per call cloning 10^5 Python objects
            t = Topology()
            t._atoms = list(self._atoms)
            return t

        def atoms(self):
            return iter(self._atoms)

        def getNumAtoms(self):
            return len(self._atoms)
