"""Minimal stand-in for `openmm.unit` (Quantity, units, md_unit_system, constants).

chiron's public API carries `openmm.unit.Quantity` everywhere (e.g. `chiron/states.py:35-91`,
`chiron/neighbors.py:446-459`, `chiron/potential.py:131-189`).  OpenMM is not installed in this
image, so this module provides the subset of the unit algebra the hot path touches.  If the real
`openmm.unit` is importable it is re-exported instead, so user code runs unchanged.

Conventions follow OpenMM: `dalton`/`amu` is gram/mole, so `md_unit_system`
(nm, ps, dalton, K, mole, radian) has kJ/mol as its energy unit; multiplying quantities whose
units cancel returns a plain number.
"""
try:  # pragma: no cover - not available in the build image
    from openmm.unit import *  # noqa: F401,F403
    from openmm import unit as _real_unit
    _HAVE_OPENMM = True
except Exception:  # noqa: BLE001
    _HAVE_OPENMM = False

if not _HAVE_OPENMM:
    import math
    import numpy as _np

    # base dimensions: length, time, mass, temperature, amount, angle, charge
    _NDIM = 7

    class Unit:
        __array_priority__ = 200
        __array_ufunc__ = None

        def __init__(self, dims, factor, name):
            self.dims = tuple(dims)
            self.factor = float(factor)  # value of one unit in SI (m, s, kg, K, mol, rad, C)
            self.name = name

        # -- algebra ---------------------------------------------------------------------
        def __mul__(self, other):
            if isinstance(other, Unit):
                return Unit([a + b for a, b in zip(self.dims, other.dims)],
                            self.factor * other.factor, f"{self.name}*{other.name}")
            if isinstance(other, Quantity):
                return Quantity(other._value, self * other.unit)._reduce()
            return Quantity(other, self)

        __rmul__ = lambda self, other: Quantity(other, self)  # noqa: E731

        def __truediv__(self, other):
            if isinstance(other, Unit):
                return Unit([a - b for a, b in zip(self.dims, other.dims)],
                            self.factor / other.factor, f"{self.name}/({other.name})")
            if isinstance(other, Quantity):
                return Quantity(1.0 / other._value, self / other.unit)._reduce()
            return Quantity(1.0 / other, self)

        def __rtruediv__(self, other):
            inv = Unit([-a for a in self.dims], 1.0 / self.factor, f"/({self.name})")
            if isinstance(other, Quantity):
                return Quantity(other._value, other.unit * inv)._reduce()
            return Quantity(other, inv)

        def __pow__(self, p):
            return Unit([a * p for a in self.dims], self.factor ** p, f"({self.name})**{p}")

        def sqrt(self):
            return self ** 0.5

        def is_compatible(self, other):
            return isinstance(other, Unit) and all(abs(a - b) < 1e-12 for a, b in zip(self.dims, other.dims))

        def is_dimensionless(self):
            return all(abs(a) < 1e-12 for a in self.dims)

        def conversion_factor_to(self, other):
            if not self.is_compatible(other):
                raise TypeError(f'Unit "{self.name}" is not compatible with Unit "{other.name}".')
            # 15 significant digits: removes binary noise such as 1e-10/1e-9 = 0.09999999999999999
            return float("%.15g" % (self.factor / other.factor))

        def in_unit_system(self, system):
            return system.unit_for(self.dims)

        def __eq__(self, other):
            return isinstance(other, Unit) and self.is_compatible(other) and \
                math.isclose(self.factor, other.factor, rel_tol=1e-12)

        def __hash__(self):
            return hash(self.dims)

        def __repr__(self):
            return f"Unit({self.name})"

        __str__ = lambda self: self.name  # noqa: E731

    class UnitSystem:
        def __init__(self, base_factors, name):
            self.base_factors = base_factors
            self.name = name

        def unit_for(self, dims):
            f = 1.0
            for d, b in zip(dims, self.base_factors):
                f *= b ** d
            return Unit(dims, f, f"{self.name}{tuple(dims)}")

    def _raw(v):
        return v

    class Quantity:
        __array_priority__ = 300
        __array_ufunc__ = None

        def __init__(self, value=None, unit=None):
            if isinstance(value, Quantity):
                if unit is None:
                    value, unit = value._value, value.unit
                else:
                    value = value.value_in_unit(unit)
            elif isinstance(value, (list, tuple)) and len(value) and isinstance(value[0], Quantity):
                u0 = value[0].unit
                value = _np.array([q.value_in_unit(u0) for q in value])
                if unit is None:
                    unit = u0
                else:
                    value = value * u0.conversion_factor_to(unit)
            if unit is None:
                unit = dimensionless
            self._value = value
            self.unit = unit

        # -- conversion -------------------------------------------------------------------
        def value_in_unit(self, unit):
            f = self.unit.conversion_factor_to(unit)
            if f == 1.0:
                return self._value
            v = _np.asarray(self._value) if isinstance(self._value, (list, tuple)) else self._value
            return v * f

        def in_units_of(self, unit):
            return Quantity(self.value_in_unit(unit), unit)

        def value_in_unit_system(self, system):
            return self.value_in_unit(self.unit.in_unit_system(system))

        def _reduce(self):
            if self.unit.is_dimensionless():
                f = self.unit.factor
                return self._value if f == 1.0 else self._value * f
            return self

        # -- arithmetic -------------------------------------------------------------------
        def __add__(self, other):
            if not isinstance(other, Quantity):
                raise TypeError("Cannot add a Quantity and a non-Quantity")
            return Quantity(self._value + other.value_in_unit(self.unit), self.unit)

        def __sub__(self, other):
            if not isinstance(other, Quantity):
                raise TypeError("Cannot subtract a Quantity and a non-Quantity")
            return Quantity(self._value - other.value_in_unit(self.unit), self.unit)

        def __neg__(self):
            return Quantity(-self._value, self.unit)

        def __mul__(self, other):
            if isinstance(other, Unit):
                return Quantity(self._value, self.unit * other)._reduce()
            if isinstance(other, Quantity):
                return Quantity(self._value * other._value, self.unit * other.unit)._reduce()
            return Quantity(self._value * other, self.unit)

        def __rmul__(self, other):
            if isinstance(other, Unit):
                return Quantity(self._value, other * self.unit)._reduce()
            return Quantity(other * self._value, self.unit)

        def __truediv__(self, other):
            if isinstance(other, Unit):
                return Quantity(self._value, self.unit / other)._reduce()
            if isinstance(other, Quantity):
                return Quantity(self._value / other._value, self.unit / other.unit)._reduce()
            return Quantity(self._value / other, self.unit)

        def __rtruediv__(self, other):
            inv = Unit([-a for a in self.unit.dims], 1.0 / self.unit.factor, f"/({self.unit.name})")
            if isinstance(other, Unit):
                return Quantity(1.0 / self._value, other * inv)._reduce()
            return Quantity(other / self._value, inv)

        def __pow__(self, p):
            return Quantity(self._value ** p, self.unit ** p)

        def sqrt(self):
            return Quantity(self._value ** 0.5, self.unit ** 0.5)

        def __imul__(self, other):
            r = self * other
            if isinstance(r, Quantity):
                self._value, self.unit = r._value, r.unit
                return self
            return r

        def __itruediv__(self, other):
            r = self / other
            if isinstance(r, Quantity):
                self._value, self.unit = r._value, r.unit
                return self
            return r

        # -- comparisons -------------------------------------------------------------------
        def _cmp_value(self, other):
            if not isinstance(other, Quantity) or not self.unit.is_compatible(other.unit):
                raise TypeError("Cannot compare quantities with incompatible units")
            return other.value_in_unit(self.unit)

        def __eq__(self, other):
            if not isinstance(other, Quantity) or not self.unit.is_compatible(other.unit):
                return False
            r = self._value == other.value_in_unit(self.unit)
            return r

        def __ne__(self, other):
            r = self.__eq__(other)
            return ~r if hasattr(r, "__invert__") and not isinstance(r, bool) else not r

        def __lt__(self, other):
            return self._value < self._cmp_value(other)

        def __le__(self, other):
            return self._value <= self._cmp_value(other)

        def __gt__(self, other):
            return self._value > self._cmp_value(other)

        def __ge__(self, other):
            return self._value >= self._cmp_value(other)

        def __bool__(self):
            return bool(self._value)

        def __hash__(self):
            return id(self)

        # -- container behaviour ---------------------------------------------------------------
        @property
        def shape(self):
            return tuple(self._value.shape)

        def __len__(self):
            return len(self._value)

        def __getitem__(self, key):
            return Quantity(self._value[key], self.unit)

        def __iter__(self):
            for v in self._value:
                yield Quantity(v, self.unit)

        def __float__(self):
            return float(self._value)

        def __repr__(self):
            return f"Quantity(value={self._value!r}, unit={self.unit.name})"

        def __str__(self):
            return f"{self._value} {self.unit.name}"

        def __format__(self, spec):
            return f"{format(self._value, spec)} {self.unit.name}"

    def is_quantity(x):
        return isinstance(x, Quantity)

    def _u(L=0, T=0, M=0, K=0, N=0, A=0, Q=0, f=1.0, name=""):
        return Unit((L, T, M, K, N, A, Q), f, name)

    dimensionless = _u(name="dimensionless")
    # length
    meter = meters = _u(L=1, f=1.0, name="m")
    nanometer = nanometers = _u(L=1, f=1e-9, name="nm")
    angstrom = angstroms = _u(L=1, f=1e-10, name="A")
    picometer = picometers = _u(L=1, f=1e-12, name="pm")
    centimeter = centimeters = _u(L=1, f=1e-2, name="cm")
    # time
    second = seconds = _u(T=1, f=1.0, name="s")
    picosecond = picoseconds = _u(T=1, f=1e-12, name="ps")
    femtosecond = femtoseconds = _u(T=1, f=1e-15, name="fs")
    nanosecond = nanoseconds = _u(T=1, f=1e-9, name="ns")
    # amount / mass (OpenMM: dalton == gram/mole)
    mole = moles = _u(N=1, f=1.0, name="mol")
    kilogram = kilograms = _u(M=1, f=1.0, name="kg")
    gram = grams = _u(M=1, f=1e-3, name="g")
    dalton = daltons = amu = amus = _u(M=1, N=-1, f=1e-3, name="Da")
    # temperature, angle, charge
    kelvin = kelvins = _u(K=1, f=1.0, name="K")
    radian = radians = _u(A=1, f=1.0, name="rad")
    degree = degrees = _u(A=1, f=math.pi / 180.0, name="deg")
    coulomb = coulombs = _u(Q=1, f=1.0, name="C")
    elementary_charge = elementary_charges = _u(Q=1, f=1.602176634e-19, name="e")
    # energy
    joule = joules = _u(L=2, T=-2, M=1, f=1.0, name="J")
    kilojoule = kilojoules = _u(L=2, T=-2, M=1, f=1e3, name="kJ")
    kilocalorie = kilocalories = _u(L=2, T=-2, M=1, f=4184.0, name="kcal")
    kilojoule_per_mole = kilojoules_per_mole = _u(L=2, T=-2, M=1, N=-1, f=1e3, name="kJ/mol")
    kilocalorie_per_mole = kilocalories_per_mole = _u(L=2, T=-2, M=1, N=-1, f=4184.0, name="kcal/mol")
    # pressure
    pascal = pascals = _u(L=-1, T=-2, M=1, f=1.0, name="Pa")
    bar = bars = _u(L=-1, T=-2, M=1, f=1e5, name="bar")
    atmosphere = atmospheres = _u(L=-1, T=-2, M=1, f=101325.0, name="atm")
    # volume
    liter = liters = litre = litres = _u(L=3, f=1e-3, name="L")

    md_unit_system = UnitSystem((1e-9, 1e-12, 1e-3, 1.0, 1.0, 1.0, 1.602176634e-19), "md")
    si_unit_system = UnitSystem((1.0,) * _NDIM, "SI")

    BOLTZMANN_CONSTANT_kB = Quantity(1.380649e-23, joule / kelvin)
    AVOGADRO_CONSTANT_NA = Quantity(6.02214076e23, dimensionless / mole)
    MOLAR_GAS_CONSTANT_R = BOLTZMANN_CONSTANT_kB * AVOGADRO_CONSTANT_NA
