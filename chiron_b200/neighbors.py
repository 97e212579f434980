"""Pair finding with chiron's interface (`chiron/neighbors.py`), computed by libchiron_b200.

`Space.displacement/wrap`, `NeighborListNsqrd` and `PairListNsqrd` keep the reference's names,
signatures, attributes, return tuples and exceptions.  Arrays are CUDA tensors (ids are int32
tensors holding the reference's uint32 values).  `NeighborListNsqrd(..., builder="cell")` selects the
O(N) counting-sort cell list; it produces the same arrays as the O(N^2) builder.
"""
import ctypes as C
from abc import ABC, abstractmethod
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib, unit
from .states import SamplerState


def _box_lengths(box_vectors) -> Tuple[float, float, float]:
    if isinstance(box_vectors, torch.Tensor):
        b = box_vectors.detach().cpu().numpy()
    else:
        b = np.asarray(box_vectors, dtype=np.float32)
    if b.shape != (3, 3):
        raise ValueError(f"box_vectors should be a 3x3 array, shape provided: {b.shape}")
    b = b.astype(np.float32)
    return float(b[0, 0]), float(b[1, 1]), float(b[2, 2])


class Space(ABC):
    """How to measure displacements and apply boundary conditions (`neighbors.py:15-36`)."""

    periodic = False

    @abstractmethod
    def displacement(self, xyz_1, xyz_2, box_vectors):
        pass

    @abstractmethod
    def wrap(self, xyz, box_vectors):
        pass

    def _displacement(self, xyz_1, xyz_2, box_vectors):
        a = _lib.as_device_f32(xyz_1)
        b = _lib.as_device_f32(xyz_2, a.device)
        a, b = torch.broadcast_tensors(a, b)
        shape = a.shape
        if len(shape) == 0 or shape[-1] != 3:
            raise ValueError(f"points must have a trailing dimension of 3, got {tuple(shape)}")
        a = a.contiguous().view(-1, 3)
        b = b.contiguous().view(-1, 3)
        lx, ly, lz = _box_lengths(box_vectors) if box_vectors is not None else (1.0, 1.0, 1.0)
        r = torch.empty_like(a)
        d = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
        _lib.get_context(a.device).call(
            "chx_displacement", _lib.ptr(a), _lib.ptr(b), a.shape[0], lx, ly, lz,
            int(self.periodic), _lib.ptr(r), _lib.ptr(d))
        return r.view(shape), d.view(shape[:-1])


class OrthogonalPeriodicSpace(Space):
    """Minimum image in an orthorhombic box (`neighbors.py:39-112`)."""

    periodic = True

    def displacement(self, xyz_1, xyz_2, box_vectors):
        if box_vectors is None:
            raise ValueError("box_vectors must be provided for a periodic system")
        return self._displacement(xyz_1, xyz_2, box_vectors)

    def wrap(self, xyz, box_vectors):
        if box_vectors is None:
            raise ValueError("box_vectors must be provided for a periodic system")
        x = _lib.as_device_f32(xyz)
        shape = x.shape
        flat = x.contiguous().view(-1, 3)
        out = torch.empty_like(flat)
        lx, ly, lz = _box_lengths(box_vectors)
        _lib.get_context(x.device).call("chx_wrap", _lib.ptr(flat), flat.shape[0], lx, ly, lz, 1,
                                        _lib.ptr(out))
        return out.view(shape)


class OrthogonalNonPeriodicSpace(Space):
    """Plain differences, identity wrap (`neighbors.py:115-175`)."""

    periodic = False

    def displacement(self, xyz_1, xyz_2, box_vectors):
        return self._displacement(xyz_1, xyz_2, box_vectors)

    def wrap(self, xyz, box_vectors):
        return _lib.as_device_f32(xyz)


def _strip_positions(positions):
    if isinstance(positions, unit.Quantity):
        if not positions.unit.is_compatible(unit.nanometer):
            raise ValueError(f"Positions require distance units, not {positions.unit}")
        positions = positions.value_in_unit_system(unit.md_unit_system)
    positions = _lib.as_device_f32(positions)
    if positions.dim() != 2 or positions.shape[1] != 3:
        raise ValueError(f"positions should be a Nx3 array, shape provided: {tuple(positions.shape)}")
    return positions


def _strip_box(box_vectors):
    if box_vectors is None:
        return None
    if isinstance(box_vectors, unit.Quantity):
        if not box_vectors.unit.is_compatible(unit.nanometer):
            raise ValueError(f"Box vectors require distance unit, not {box_vectors.unit}")
        box_vectors = box_vectors.value_in_unit_system(unit.md_unit_system)
    shape = tuple(np.shape(box_vectors)) if not isinstance(box_vectors, torch.Tensor) else tuple(box_vectors.shape)
    if shape != (3, 3):
        raise ValueError(f"box_vectors should be a 3x3 array, shape provided: {shape}")
    return _lib.as_device_f32(box_vectors)


class PairsBase(ABC):
    """Common constructor / validation / `build_from_state` (`neighbors.py:178-375`)."""

    def __init__(self, space: Space, cutoff: Optional[unit.Quantity] = unit.Quantity(1.2, unit.nanometer)):
        if not isinstance(space, Space):
            raise TypeError(f"space must be of type Space, found {type(space)}")
        if not cutoff.unit.is_compatible(unit.angstrom):
            raise ValueError(
                f"cutoff must be a unit.Quantity with units of distance, cutoff.unit = {cutoff.unit}")
        self.cutoff = cutoff
        self.space = space

    @abstractmethod
    def build(self, positions, box_vectors):
        pass

    def _validate_build_inputs(self, positions, box_vectors):
        self.ref_positions = _strip_positions(positions)
        self.box_vectors = _strip_box(box_vectors)

    def build_from_state(self, sampler_state: SamplerState):
        if not isinstance(sampler_state, SamplerState):
            raise TypeError(f"Expected SamplerState, got {type(sampler_state)} instead")
        self.build(sampler_state.positions, sampler_state.box_vectors)

    @abstractmethod
    def calculate(self, positions):
        pass

    @abstractmethod
    def check(self, positions) -> bool:
        pass

    def _box_args(self):
        if self.space.periodic:
            if self.box_vectors is None:
                raise ValueError("box_vectors must be provided for a periodic system")
            return _box_lengths(self.box_vectors) + (1,)
        return (1.0, 1.0, 1.0, 0)


class NeighborListNsqrd(PairsBase):
    """Half Verlet list with skin, padded to `n_max_neighbors` (`neighbors.py:378-907`).

    Semantics follow the reference, with one documented fix: the list grows
    whenever a row holds `>= n_max_neighbors` entries (the reference only reacts to `==` and can
    silently truncate).  `builder` is "nsq" (the reference's O(N^2) algorithm) or "cell".
    """

    def __init__(self, space: Space, cutoff: unit.Quantity = unit.Quantity(1.2, unit.nanometer),
                 skin: unit.Quantity = unit.Quantity(0.4, unit.nanometer), n_max_neighbors: float = 200,
                 builder: str = "nsq"):
        if not isinstance(space, Space):
            raise TypeError(f"space must be of type Space, found {type(space)}")
        if not skin.unit.is_compatible(unit.angstrom):
            raise ValueError(f"cutoff must be a unit.Quantity with units of distance, skin.unit = {skin.unit}")
        if builder not in ("nsq", "cell"):
            raise ValueError(f"builder must be 'nsq' or 'cell', got {builder!r}")
        self.cutoff = cutoff
        self.skin = skin
        self.n_max_neighbors = n_max_neighbors
        self.space = space
        self.builder = builder
        self.is_built = False
        self.n_builds = 0

    def __deepcopy__(self, memo):
        """`MCMCSampler.run` deep-copies its inputs (mcmc.py:1134-1136) so that the caller's objects are
        never touched.  The list arrays are immutable here -- `build` replaces them wholesale and the
        device loops write into their own copies -- so a deep copy shares the (N, M) tensors instead of
        moving 2 x N x M words per call; everything else is copied."""
        import copy as _copy
        new = _copy.copy(self)
        shared = ("_arrays", "ref_positions", "box_vectors", "particle_ids")
        for k, v in self.__dict__.items():
            if k not in shared:
                new.__dict__[k] = _copy.deepcopy(v, memo)
        return new

    @property
    def cutoff(self) -> unit.Quantity:
        return self._cutoff

    @cutoff.setter
    def cutoff(self, cutoff: unit.Quantity) -> None:
        if not cutoff.unit.is_compatible(unit.nanometer):
            raise ValueError(f"cutoff must be a unit.Quantity with units of distance, cutoff.unit = {cutoff.unit}")
        self._cutoff = cutoff
        self.is_built = False

    @property
    def skin(self) -> unit.Quantity:
        return self._skin

    @skin.setter
    def skin(self, skin: unit.Quantity) -> None:
        if not skin.unit.is_compatible(unit.nanometer):
            raise ValueError(f"skin must be a unit.Quantity with units of distance, skin.unit = {skin.unit}")
        self._skin = skin
        self.is_built = False

    # unit-free scalars used by the kernels
    def _cutoff_md(self) -> float:
        return float(self.cutoff.value_in_unit_system(unit.md_unit_system))

    def _skin_md(self) -> float:
        return float(self.skin.value_in_unit_system(unit.md_unit_system))

    def _cutoff_plus_skin_md(self) -> float:
        return float((self.cutoff + self.skin).value_in_unit_system(unit.md_unit_system))

    # The padded arrays are produced on demand: the fused Langevin engine tracks the reference's
    # rebuild events on the device and only hands back the reference positions of the last one
    # (`_adopt_reference`); the (N, n_max) arrays are then materialised when somebody reads them.
    def _adopt_reference(self, ref_positions, box_vectors, n_events=1):
        self.ref_positions = ref_positions
        self.box_vectors = _strip_box(box_vectors)
        self._arrays = None
        self.is_built = True
        self.n_builds += int(n_events)

    def _materialise(self):
        if getattr(self, "_arrays", None) is None:
            self._build_arrays(self.ref_positions)
        return self._arrays

    @property
    def neighbor_list(self):
        return self._materialise()[0]

    @property
    def neighbor_mask(self):
        return self._materialise()[1]

    @property
    def n_neighbors(self):
        return self._materialise()[2]

    def build(self, positions, box_vectors):
        positions = _strip_positions(positions)
        box_vectors = _strip_box(box_vectors)
        if self.space.periodic and box_vectors is None:
            raise ValueError("box_vectors must be provided for a periodic system")
        self.ref_positions = positions
        self.box_vectors = box_vectors
        self._build_arrays(positions)
        self.is_built = True
        self.n_builds += 1

    def _build_arrays(self, positions):
        n = positions.shape[0]
        dev = positions.device
        self.particle_ids = torch.arange(n, dtype=torch.int32, device=dev)
        lx, ly, lz, periodic = self._box_args()
        ctx = _lib.get_context(dev)
        fn = "chx_nlist_build_cell" if self.builder == "cell" else "chx_nlist_build_nsq"
        max_count, n_eq = C.c_int(0), C.c_int(0)
        M = int(self.n_max_neighbors)
        while True:
            nl = torch.empty((n, M), dtype=torch.int32, device=dev)
            mask = torch.empty((n, M), dtype=torch.int32, device=dev)
            nn = torch.empty((n,), dtype=torch.int32, device=dev)
            ctx.call(fn, _lib.ptr(positions), n, lx, ly, lz, periodic, self._cutoff_plus_skin_md(), M,
                     _lib.ptr(nl), _lib.ptr(mask), _lib.ptr(nn), C.byref(max_count), C.byref(n_eq))
            if max_count.value >= M:
                from loguru import logger as log
                log.debug(f"Increasing n_max_neighbors from {M} to at  {max_count.value + 10}")
                M = max_count.value + 10
                continue
            break
        self.n_max_neighbors = M
        self._arrays = (nl, mask, nn)

    def calculate(self, positions):
        positions = _lib.as_device_f32(positions)
        n, M = self.neighbor_list.shape
        dev = positions.device
        n_out = torch.empty((n,), dtype=torch.int32, device=dev)
        mask = torch.empty((n, M), dtype=torch.int32, device=dev)
        dist = torch.empty((n, M), dtype=torch.float32, device=dev)
        rij = torch.empty((n, M, 3), dtype=torch.float32, device=dev)
        lx, ly, lz, periodic = self._box_args()
        _lib.get_context(dev).call(
            "chx_nlist_calculate", _lib.ptr(positions), n, lx, ly, lz, periodic, self._cutoff_md(), M,
            _lib.ptr(self.neighbor_list), _lib.ptr(self.neighbor_mask), _lib.ptr(n_out), _lib.ptr(mask),
            _lib.ptr(dist), _lib.ptr(rij))
        return n_out, self.neighbor_list, mask, dist, rij

    def check_async(self, positions) -> torch.Tensor:
        """Device flag (int32 scalar tensor) of the rebuild condition, no host sync."""
        positions = _lib.as_device_f32(positions)
        flag = torch.zeros((), dtype=torch.int32, device=positions.device)
        lx, ly, lz, periodic = self._box_args()
        _lib.get_context(positions.device).call(
            "chx_nlist_check", _lib.ptr(positions), _lib.ptr(self.ref_positions), positions.shape[0],
            lx, ly, lz, periodic, float(np.float32(self._skin_md() / 2.0)), _lib.ptr(flag))
        return flag

    def check(self, positions) -> bool:
        if self.ref_positions.shape[0] != positions.shape[0]:
            return True
        return bool(self.check_async(positions).item())


class PairListNsqrd(PairsBase):
    """All pairs, O(N^2) (`neighbors.py:910-1289`).  `all_pairs` / `reduction_mask` are produced
    on demand; `calculate` and the energy kernels never materialise them."""

    def __init__(self, space: Space, cutoff: Optional[unit.Quantity] = None):
        if not isinstance(space, Space):
            raise TypeError(f"space must be of type Space, found {type(space)}")
        self.cutoff = cutoff
        self.space = space
        self.is_built = False

    @property
    def cutoff(self):
        return self._cutoff

    @cutoff.setter
    def cutoff(self, cutoff):
        if cutoff is not None and not cutoff.unit.is_compatible(unit.angstrom):
            raise ValueError(f"cutoff must be a unit.Quantity with units of distance, cutoff.unit = {cutoff.unit}")
        self._cutoff = cutoff

    def _cutoff_md(self) -> float:
        return -1.0 if self.cutoff is None else float(self.cutoff.value_in_unit_system(unit.md_unit_system))

    def build(self, positions, box_vectors):
        self._validate_build_inputs(positions, box_vectors)
        if self.space.periodic and self.box_vectors is None:
            raise ValueError("box_vectors must be provided for a periodic system")
        self.n_particles = self.ref_positions.shape[0]
        self.particle_ids = torch.arange(self.n_particles, dtype=torch.int32, device=self.ref_positions.device)
        self._pairs = None
        self.is_built = True

    def _materialise(self):
        if self._pairs is None:
            n, dev = self.n_particles, self.ref_positions.device
            pairs = torch.empty((n, max(n - 1, 0)), dtype=torch.int32, device=dev)
            red = torch.empty((n, max(n - 1, 0)), dtype=torch.uint8, device=dev)
            if n > 1:
                _lib.get_context(dev).call("chx_pairlist_build", n, _lib.ptr(pairs), _lib.ptr(red))
            self._pairs = (pairs, red.bool())
        return self._pairs

    @property
    def all_pairs(self):
        return self._materialise()[0]

    @property
    def reduction_mask(self):
        return self._materialise()[1]

    def calculate(self, positions):
        positions = _lib.as_device_f32(positions)
        if positions.shape[0] != self.n_particles:
            raise ValueError(
                f"Number of particles cannot changes without rebuilding. "
                f"Positions must have shape ({self.n_particles}, 3), found {tuple(positions.shape)}")
        n, dev = self.n_particles, positions.device
        n_out = torch.zeros((n,), dtype=torch.int32, device=dev)
        mask = torch.empty((n, n - 1), dtype=torch.int32, device=dev)
        dist = torch.empty((n, n - 1), dtype=torch.float32, device=dev)
        rij = torch.empty((n, n - 1, 3), dtype=torch.float32, device=dev)
        if n > 1:
            lx, ly, lz, periodic = self._box_args()
            _lib.get_context(dev).call(
                "chx_pairlist_calculate", _lib.ptr(positions), n, lx, ly, lz, periodic, self._cutoff_md(),
                _lib.ptr(n_out), _lib.ptr(mask), _lib.ptr(dist), _lib.ptr(rij))
        return n_out, self.all_pairs, mask, dist, rij

    def check(self, positions) -> bool:
        return positions.shape[0] != self.n_particles
