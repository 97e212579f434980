"""Reporter surface of `chiron/reporters.py` backed by memory / `.npz` files.

HDF5 (h5py) and XTC (mdtraj) are not available in this image and file formats are outside the
hot path; the classes keep the calls the integrator, the moves and the
multistate sampler make: `report(dict)`, `flush_buffer()`, `get_property(name)`,
`get_available_keys()`, `reset_reporter_file()`, `log_file_path`, `xtc_file_path`, `workdir`.
Device tensors are moved to the host lazily (at flush / get_property), never per report.
"""
import os
from typing import List, Optional

import numpy as np
import torch


def _host(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


class BaseReporter:
    _directory = None

    @classmethod
    def set_directory(cls, directory: str):
        cls._directory = str(directory)

    @classmethod
    def get_directory(cls):
        from pathlib import Path
        if cls._directory is None:
            cls._directory = str(Path.cwd() / "chiron_output")
        os.makedirs(cls._directory, exist_ok=True)
        return cls._directory


class _SimulationReporter:
    _default_properties: List[str] = []

    def __init__(self, file_name: str, buffer_size: int = 10, default_properties=None):
        if default_properties is not None:
            self._default_properties = list(default_properties)
        self._file_name = file_name
        self.buffer_size = buffer_size
        self._buffer = {}
        self._store = {}

    # paths -------------------------------------------------------------------------------------------
    @property
    def workdir(self):
        return BaseReporter.get_directory()

    @property
    def log_file_path(self):
        return os.path.join(self.workdir, f"{self._file_name}.npz")

    @property
    def xtc_file_path(self):
        return os.path.join(self.workdir, f"{self._file_name}_traj.npz")

    @property
    def properties_to_report(self):
        return self._default_properties

    # reporting ---------------------------------------------------------------------------------------
    def report(self, data_dict: dict) -> None:
        for key, value in data_dict.items():
            if self._default_properties and key not in self._default_properties:
                continue
            if isinstance(value, torch.Tensor):
                value = value.detach().clone()
            self._buffer.setdefault(key, []).append(value)

    def flush_buffer(self) -> None:
        for key, values in self._buffer.items():
            self._store.setdefault(key, []).extend(_host(v) for v in values)
        self._buffer = {}
        if self._store:
            np.savez(self.log_file_path, **{k: np.asarray(v) for k, v in self._store.items()})
            if "positions" in self._store:
                np.savez(self.xtc_file_path, positions=np.asarray(self._store["positions"]))

    def reset_reporter_file(self):
        self._buffer, self._store = {}, {}
        for p in (self.log_file_path, self.xtc_file_path):
            if os.path.exists(p):
                os.remove(p)

    def get_available_keys(self):
        return sorted(set(self._store) | set(self._buffer))

    def get_property(self, name: str):
        values = list(self._store.get(name, [])) + [_host(v) for v in self._buffer.get(name, [])]
        if not values:
            return None
        return np.asarray(values)

    def close(self):
        self.flush_buffer()


class LangevinDynamicsReporter(_SimulationReporter):
    _name = "langevin_reporter"

    def __init__(self, name: str = None, buffer_size: int = 1, topology=None):
        super().__init__(file_name=name or self._name, buffer_size=buffer_size,
                         default_properties=["positions", "box_vectors", "potential_energy", "step",
                                             "iteration", "elapsed_step"])
        self.topology = topology

    @classmethod
    def get_name(cls):
        return cls._name

    def report(self, data_dict: dict) -> None:
        # the reference writes positions to an XTC trajectory and only scalars to the log
        super().report(data_dict)


class MCReporter(_SimulationReporter):
    _name = "mc_reporter"

    def __init__(self, name=None, buffer_size: int = 1):
        label = self._name if name is None or isinstance(name, int) else str(name)
        super().__init__(file_name=label, buffer_size=buffer_size, default_properties=[])

    @classmethod
    def get_name(cls):
        return cls._name


class MultistateReporter(_SimulationReporter):
    _name = "multistate_reporter"

    def __init__(self, name: str = None, buffer_size: int = 1):
        super().__init__(file_name=name or self._name, buffer_size=buffer_size,
                         default_properties=["positions", "box_vectors", "u_kn", "state_index", "step"])
        self._replica_reporter = {}

    @classmethod
    def get_name(cls):
        return cls._name
