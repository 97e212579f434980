"""Python side of the fused LJ Langevin engine (`chx_ljmd_*`, include/chiron_b200.h)."""
import ctypes as C

import numpy as np
import torch

from . import _lib, unit


class LjmdParams(C.Structure):
    _fields_ = [("n", C.c_int), ("lx", C.c_float), ("ly", C.c_float), ("lz", C.c_float),
                ("sigma", C.c_float), ("epsilon", C.c_float), ("cutoff", C.c_float),
                ("skin", C.c_float), ("dt", C.c_float), ("gamma", C.c_float), ("kT", C.c_float),
                ("n_replicas", C.c_int), ("internal_skin", C.c_float)]


_P, _I = C.c_void_p, C.c_int
_SIGS = {
    "chx_ljmd_create": [_P, C.POINTER(LjmdParams), C.POINTER(_P)],
    "chx_ljmd_destroy": [_P],
    "chx_ljmd_set_state": [_P, _P, _P, _P, C.POINTER(C.c_float)],
    "chx_ljmd_get_state": [_P, _P, _P, _P, _P],
    "chx_ljmd_run": [_P, _I, C.POINTER(C.c_uint32), _I, _P, _I],
    "chx_ljmd_energy": [_P, _P],
    "chx_ljmd_set_kt": [_P, C.POINTER(C.c_float)],
    "chx_ljmd_scale_velocities": [_P, C.POINTER(C.c_float)],
    "chx_ljmd_stats": [_P, C.POINTER(C.c_longlong)],
    "chx_ljmd_table_stats": [_P, C.POINTER(C.c_longlong)],
    "chx_ljmd_force_only": [_P, _I],
    "chx_ljmd_set_chunk_phase": [_P, _I, _I],
    "chx_ljmd_set_gpu_share": [_P, _I],
    "chx_ljmd_set_prebuild": [_P, _I],
    "chx_ljmd_prebuild_now": [_P],
    "chx_ljmd_step_timing": [_P, C.POINTER(C.c_double), C.POINTER(C.c_longlong), _I],
    "chx_fma_peak": [_P, _I, C.POINTER(C.c_double)],
}
_lib.SIGNATURES.update(_SIGS)

# engine tuning knob: skin of the internal tables (None = the list's own skin)
INTERNAL_SKIN_NM = None


def available() -> bool:
    return torch.cuda.is_available()


def box_supported(nbr_list, sampler_state) -> bool:
    return sampler_state.box_vectors is not None


class LJLangevinEngine:
    """Owns a chx_ljmd: R replicas x N particles, one box, one LJ parameter set."""

    def __init__(self, n, box, sigma, epsilon, cutoff, skin, dt, gamma, kT, n_replicas=1,
                 internal_skin=None, device=None, ctx=None):
        # ctx: a private _lib.Context (own chx_ctx and stream) for an engine driven from its own host thread
        self.ctx = ctx if ctx is not None else _lib.get_context(device)
        self.device = self.ctx.device
        self.n, self.R = int(n), int(n_replicas)
        if internal_skin is None:
            internal_skin = INTERNAL_SKIN_NM
        p = LjmdParams(int(n), float(box[0]), float(box[1]), float(box[2]), float(sigma), float(epsilon),
                       float(cutoff), float(skin), float(dt), float(gamma), float(kT), int(n_replicas),
                       float(internal_skin or 0.0))
        self.params = p
        h = _P()
        self.ctx.sync_stream()
        _lib.check(self.ctx.lib.chx_ljmd_create(self.ctx.handle, C.byref(p), C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.chx_ljmd_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        self.ctx.sync_stream()
        _lib.check(getattr(self.ctx.lib, name)(self.handle, *args))

    def set_state(self, x, v, mass, kT_per_replica=None):
        x = _lib.as_device_f32(x, self.device)
        v = _lib.as_device_f32(v, self.device)
        mass = _lib.as_device_f32(mass, self.device)
        assert x.numel() == self.R * self.n * 3 and v.numel() == x.numel() and mass.numel() == self.n
        kt = None
        if kT_per_replica is not None:
            kt = (C.c_float * self.R)(*[float(t) for t in kT_per_replica])
        self._call("chx_ljmd_set_state", _lib.ptr(x), _lib.ptr(v), _lib.ptr(mass), kt)

    def get_state(self, want_force=False, want_ref=False):
        shape = (self.R, self.n, 3) if self.R > 1 else (self.n, 3)
        x = torch.empty(shape, dtype=torch.float32, device=self.device)
        v = torch.empty(shape, dtype=torch.float32, device=self.device)
        f = torch.empty(shape, dtype=torch.float32, device=self.device) if want_force else None
        ref = torch.empty(shape, dtype=torch.float32, device=self.device) if want_ref else None
        self._call("chx_ljmd_get_state", _lib.ptr(x), _lib.ptr(v), _lib.ptr(f), _lib.ptr(ref))
        return x, v, f, ref

    def run(self, nsteps, keys, report_interval=0):
        """keys: (R,2) uint32 loop keys; returns (keys_after, energies (n_reports, R) float64 tensor or None)."""
        keys = np.ascontiguousarray(np.asarray(keys, dtype=np.uint32).reshape(self.R, 2)).copy()
        n_rep = (nsteps + report_interval - 1) // report_interval if report_interval > 0 else 0
        energies = torch.zeros((n_rep, self.R), dtype=torch.float64, device=self.device) if n_rep else None
        self._call("chx_ljmd_run", int(nsteps), keys.ctypes.data_as(C.POINTER(C.c_uint32)),
                   int(report_interval), _lib.ptr(energies), int(n_rep))
        return keys, energies

    def energy(self):
        e = torch.zeros((self.R,), dtype=torch.float64, device=self.device)
        self._call("chx_ljmd_energy", _lib.ptr(e))
        return e

    def set_kT(self, kT_per_replica):
        self._call("chx_ljmd_set_kt", (C.c_float * self.R)(*[float(t) for t in kT_per_replica]))

    def scale_velocities(self, scale_per_replica):
        self._call("chx_ljmd_scale_velocities", (C.c_float * self.R)(*[float(t) for t in scale_per_replica]))

    def set_chunk_phase(self, num, den):
        """Shift the chunk grid of batched replicas by num/den of a chunk (engines that share a GPU)."""
        self._call("chx_ljmd_set_chunk_phase", int(num), int(den))

    def set_gpu_share(self, n_engines):
        """This engine shares the GPU with n_engines - 1 others running at the same time."""
        self._call("chx_ljmd_set_gpu_share", int(n_engines))

    def set_prebuild(self, mode=1):
        """Overlap the table rebuild a run ends on with the caller's work between two runs (replica exchange):
        1 = enqueued by the run, 2 = enqueued by `prebuild_now()`, 0 = off."""
        self._call("chx_ljmd_set_prebuild", int(mode))

    def prebuild_now(self):
        self._call("chx_ljmd_prebuild_now")

    def force_only(self, repeats=1):
        self._call("chx_ljmd_force_only", int(repeats))

    def step_timing(self, reset=False):
        """(total ms, launches) of the step kernel over the chunk replays without a table rebuild."""
        ms, steps = C.c_double(0.0), C.c_longlong(0)
        self._call("chx_ljmd_step_timing", C.byref(ms), C.byref(steps), int(bool(reset)))
        return ms.value, steps.value

    def stats(self):
        s = (C.c_longlong * 8)()
        self._call("chx_ljmd_stats", s)
        keys = ("table_rebuilds", "candidate_pairs", "interacting_pairs", "steps", "launches",
                "reference_rebuilds", "tile_capacity", "blocks")
        out = dict(zip(keys, [int(v) for v in s]))
        t = (C.c_longlong * 4)()
        self._call("chx_ljmd_table_stats", t)
        out.update(trip_slots=int(t[0]), list_words=int(t[1]), tile_bytes=int(t[2]), candidate_capacity=int(t[3]))
        out["lane_utilisation"] = (2.0 * out["candidate_pairs"] / out["trip_slots"]) if out["trip_slots"] else 0.0
        return out


def run_fused_langevin(integrator, x, v, mass, potential, nbr_list, sampler_state, kT, dt, gamma, key,
                       number_of_steps):
    """LangevinIntegrator.run body for LJ + NeighborListNsqrd + periodic box on the fused engine."""
    box = sampler_state.box_lengths_host()
    n = x.shape[0]
    sig = (n, box, potential.sigma, potential.epsilon, potential.cutoff, nbr_list._skin_md(), dt, gamma)
    eng = getattr(integrator, "_engine", None)
    if eng is None or getattr(integrator, "_engine_sig", None) != sig or eng.device != x.device:
        if eng is not None:
            eng.close()
        eng = LJLangevinEngine(n, box, potential.sigma, potential.epsilon, potential.cutoff,
                               nbr_list._skin_md(), dt, gamma, kT, device=x.device)
        integrator._engine, integrator._engine_sig = eng, sig
    eng.set_state(x, v, mass, [kT])
    has_reporter = getattr(integrator, "reporter", None) is not None
    interval = int(integrator.report_interval)
    want_traj = has_reporter or integrator.save_traj_in_memory
    before = eng.stats()
    if want_traj:
        # positions are reported too: run report_interval steps at a time
        # (step s reports after s+1 steps, integrators.py:197-205)
        done = 0
        keys = np.asarray(key, dtype=np.uint32).reshape(1, 2)
        while done < number_of_steps:
            nxt = done if done % interval == 0 else (done // interval + 1) * interval
            seg = min(number_of_steps, nxt + 1) - done
            keys, _ = eng.run(seg, keys, 0)
            done += seg
            if (done - 1) % interval == 0:
                xs, _, _, _ = eng.get_state()
                energy = eng.energy()[0].float()
                step = done - 1
                if has_reporter:
                    integrator._report(xs, potential, nbr_list, step, integrator._move_iteration,
                                       step + integrator._move_iteration * number_of_steps, energy=energy)
                if integrator.save_traj_in_memory:
                    integrator.traj.append(xs)
        key = keys[0]
    else:
        keys, _ = eng.run(number_of_steps, np.asarray(key, dtype=np.uint32).reshape(1, 2), 0)
        key = keys[0]
    x, v, _, ref = eng.get_state(want_ref=True)
    after = eng.stats()
    # hand the reference-rebuild bookkeeping back to the list object (arrays rebuilt lazily)
    if after["reference_rebuilds"] > before["reference_rebuilds"]:
        nbr_list._adopt_reference(ref, sampler_state.box_vectors,
                                  after["reference_rebuilds"] - before["reference_rebuilds"])
    integrator.last_run_stats = {"path": "fused", **after}
    return x, v, key
