#!/usr/bin/env python
"""Benchmark of the particle hot path: LJ argon fluid, N = 262,144, rho* = 0.8, Langevin BAOAB with
the cell-list neighbour build (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--inner S] [--impl ours|reference]

A bench "step" is one `LangevinIntegrator.run`-sized chunk of S = --inner BAOAB steps (default 1000, the
call Examples/LJ_langevin.py makes);
`value` is BAOAB steps/s summed over all ranks (one independent system per GPU: the single-system
path does not shard, SURVEY.md section 8e -> weak scaling, "replicas only").
  value  device-resident: state stays in HBM, K chunks timed with CUDA events on the launch stream
  e2e    through the public API with HOST buffers: every chunk uploads positions/velocities from
         pinned host memory, runs S steps and reads positions, velocities and the energy back
  roofline   FP32 CUDA-core roofline of the dominant kernel (k_md_force), algorithmic FLOPs
             18 P_cand + 21 P_int per launch / CUDA-event duration, against an
             FFMA-chain peak measured in the same run; HBM figures beside it
  cpu_baseline   the oracle (NumPy restatement of the reference algorithm) timed on a bounded sample
`--impl reference` times the reference algorithm's CPU restatement (JAX cannot be installed here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIGMA, EPS_KCAL, RC, SKIN = 0.34, 0.238, 1.02, 0.5
EPS = EPS_KCAL * 4.184
MASS, TEMP_K, DT_PS, GAMMA = 39.948, 300.0, 0.001, 1.0
N_SIDE, RHO_STAR = 64, 0.8
# dram bytes of one step-kernel launch: read at bench time from the committed ncu export of the same kernel
NCU_STEP_KERNEL_CSV = os.path.join(ROOT, "profiles", "r02_step_kernel_raw.csv")


def ncu_step_kernel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of k_md_force<UPDATE> from the committed
    `ncu --set full --clock-control none --cache-control none` export (warm L2: the launch sits inside the step
    loop, tables and state L2 resident); None if the file is missing."""
    import csv
    try:
        tot, unit_scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        with open(NCU_STEP_KERNEL_CSV) as fh:
            for row in csv.DictReader(fh):
                if row["metric"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(row["value"]) * unit_scale.get(row["unit"], 1.0)
        return tot or None
    except Exception:
        return None
WORKLOAD = "LJ argon fluid N=262144 rho*=0.8 rc=3sigma skin=0.5nm T=300K dt=1fs Langevin BAOAB, cell-list build"


def bench_config(n_side):
    """`config` of the JSON line: the workload only, identical in both arms (--impl ours / reference)."""
    n = n_side ** 3
    return {"workload": WORKLOAD if n_side == N_SIDE else WORKLOAD.replace("262144", str(n)), "n_particles": n,
            "parallelism": "1 independent system per GPU (replicas only)"}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--inner", type=int, default=1000,
                   help="BAOAB steps per bench step = per LangevinIntegrator.run call (Examples/LJ_langevin.py runs 1000)")
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n-side", type=int, default=N_SIDE)
    p.add_argument("--internal-skin", type=float, default=None)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-remd", action="store_true")
    p.add_argument("--no-mc", action="store_true", help="skip the Monte Carlo configs (secondary numbers)")
    p.add_argument("--remd-sweeps", type=int, default=20)
    return p.parse_args()


def make_system(n_side, seed):
    from chiron_b200 import unit
    from chiron_b200.testsystems import LennardJonesFluid
    lj = LennardJonesFluid(nparticles=n_side ** 3, reduced_density=RHO_STAR, sigma=SIGMA * unit.nanometer,
                           epsilon=EPS_KCAL * unit.kilocalories_per_mole, seed=seed)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=np.float32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=np.float32)
    return lj, x, box


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region: NVML in-process (initialised before the
    region starts, one sample every 10 ms), `nvidia-smi` as a fallback."""

    _Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
          "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torchrun / CUDA_VISIBLE_DEVICES: NVML counts physical devices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.nvml = (pynvml, h, mx)
        except Exception:
            self.nvml = None

    def _nvml_sample(self):
        pynvml, h, mx = self.nvml
        bits = (pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap)
        sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        try:
            power = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        except Exception:
            power = float("nan")
        reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        self.rows.append([str(sm), str(mx), str(power)] + ["Active" if reasons & b else "Not Active" for b in bits])

    def _smi_sample(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self._Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout
        row = [c.strip() for c in out.strip().split(",")]
        if len(row) >= 7:
            self.rows.append(row)

    def _run(self):
        while not self.stop.is_set():
            try:
                if self.nvml is not None:
                    self._nvml_sample()
                else:
                    self._smi_sample()
            except Exception:
                self.nvml = None
            self.stop.wait(float(os.environ.get("CHX_BENCH_CLOCK_S", "0.01")) if self.nvml is not None else 0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=12)
        if not self.rows:      # nothing during the region: one sample right behind it is better than none
            for fn in (self._nvml_sample if self.nvml is not None else None, self._smi_sample):
                try:
                    if fn is not None and not self.rows:
                        fn()
                except Exception:
                    pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, val in zip(names, r[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        power = []
        for r in self.rows:
            try:
                power.append(float(r[2]))
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": (max(power) if power else None)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the C/OpenMP restatement of the reference algorithm (oracle/c), all host threads
# ---------------------------------------------------------------------------------------------------
REF_STEPS_PER_BENCH_STEP = 20      # full-size BAOAB steps per bench step of the reference arm
REF_INTERVAL_PROBE_STEPS = 450     # trajectory length on which the reference arm measures its own rebuild interval


def cpu_probe_rebuild_interval(x, box, v0, nsteps=REF_INTERVAL_PROBE_STEPS):
    """Rebuild interval of the REFERENCE list (check(): any particle skin/2 from its reference position,
    neighbors.py:864-907) on a trajectory of `nsteps` full-size steps of the C port, started from the bench's
    initial state.  The rebuilds of the probe use the oracle's cell-grid accelerator and are not timed."""
    from oracle import cport
    n = x.shape[0]
    cport.set_build_mode(1)
    try:
        xo, vo, key, st = cport.langevin_lj(x, v0, np.full(n, MASS, np.float32), box, SIGMA, EPS, RC, SKIN, 400,
                                            8.314462618e-3 * TEMP_K, DT_PS, GAMMA, np.array([0, 1234], np.uint32), nsteps)
    finally:
        cport.set_build_mode(0)
    events = int(st["n_builds"]) - 1
    return (nsteps / events if events > 0 else None), events, (xo, vo, key)


def cpu_reference_bench_steps(x, box, v0, k_steps, warmup, interval, n_s=REF_STEPS_PER_BENCH_STEP, state=None):
    """K bench steps of the reference algorithm on the host cores.  One bench step is executed, not modelled:
      * `n_s` full-size BAOAB steps of a running trajectory over the padded (N, M) reference list (calculate +
        masked LJ energy + the autodiff-equivalent force + BAOAB with in-stream threefry noise + check), timed as
        the difference between an (n_s + 2)-step and a 2-step call from the same state (so the force evaluation
        that opens a call and the untimed set-up build cancel);
      * the rows i = k (mod stride) of the reference's O(N^2) list build on the current positions, stride =
        round(interval / n_s): over `stride` bench steps every row of one complete build has been executed, at the
        rate at which the reference rebuilds (every `interval` steps).
    Returns (BAOAB steps/s, detail)."""
    from oracle import cport
    cport.use_all_cores()
    n = x.shape[0]
    kT = 8.314462618e-3 * TEMP_K
    mass = np.full(n, MASS, np.float32)
    if state is None:
        state = (x, v0, np.array([0, 1234], np.uint32))
    xs, vs, key = state
    stride = max(1, int(round(interval / n_s))) if interval else 0
    times, t_steps, t_rows_l = [], [], []
    M = 400
    for k in range(warmup + k_steps):
        cport.set_build_mode(1)
        try:
            _, _, _, st_a = cport.langevin_lj(xs, vs, mass, box, SIGMA, EPS, RC, SKIN, M, kT, DT_PS, GAMMA, key, 2)
            xs2, vs2, key2, st_b = cport.langevin_lj(xs, vs, mass, box, SIGMA, EPS, RC, SKIN, M, kT, DT_PS, GAMMA, key, 2 + n_s)
        finally:
            cport.set_build_mode(0)
        t_s = max(1e-9, st_b["t_steps_s"] - st_a["t_steps_s"])
        t_r = 0.0
        if stride:
            t_r, _ = cport.time_reference_build_rows(xs2, box, np.float32(RC + SKIN), stride, row0=k % stride)
        xs, vs, key = xs2, vs2, key2
        M = int(st_b["M"])
        if k >= warmup:
            times.append(t_s + t_r); t_steps.append(t_s); t_rows_l.append(t_r)
    total = float(np.sum(times))
    detail = {"bench_steps_timed": k_steps, "baoab_steps_per_bench_step": n_s, "build_row_stride": stride,
              "rebuild_interval_steps": interval, "t_steps_s_per_bench_step": float(np.mean(t_steps)),
              "t_build_rows_s_per_bench_step": float(np.mean(t_rows_l)),
              "t_full_build_s": float(np.mean(t_rows_l)) * stride if stride else None,
              "full_builds_executed": (k_steps / stride) if stride else 0.0,
              "n_max_neighbors": M, "p_cand": int(st_b["p_cand"]), "threads": cport.num_threads(), "extrapolated": False}
    return k_steps * n_s / total, total / k_steps, detail, (xs, vs, key)


CPU_SAMPLE = ("%d bench steps EXECUTED at full size (N=%d): each = %d BAOAB steps of a running trajectory over the padded "
              "(N,M) reference list (calculate + masked LJ energy/force + threefry noise + check) + the rows i = k mod %d of "
              "the reference's O(N^2) build (one complete build per %d bench steps = the measured rebuild interval of %s "
              "steps); C/OpenMP restatement of the reference algorithm (oracle/c) on all host threads -- chiron's JAX path "
              "cannot run this size (N x N mask) and jax is not installable offline")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge
    ge.build_oracle()
    from oracle import cport, jax_random as jr, dynamics as dyn
    cport.use_all_cores()
    lj, x, box = make_system(args.n_side, seed=4)
    n = x.shape[0]
    v0 = dyn.maxwell_boltzmann(jr.PRNGKey(11), np.full(n, MASS), TEMP_K)
    t0 = time.perf_counter()
    interval, events, state = cpu_probe_rebuild_interval(x, box, v0)
    t_probe = time.perf_counter() - t0
    value, t_bench_step, detail, _ = cpu_reference_bench_steps(x, box, v0, args.steps, args.warmup, interval, state=state)
    detail.update(interval_probe_steps=REF_INTERVAL_PROBE_STEPS, interval_probe_events=events, interval_probe_wall_s=t_probe)
    sample = CPU_SAMPLE % (args.steps, n, detail["baoab_steps_per_bench_step"], max(1, detail["build_row_stride"]),
                           max(1, detail["build_row_stride"]), ("%.0f" % interval) if interval else "no event in the probe: never")
    line = {
        "impl": "reference", "metric": "LJ Langevin steps/s at N=262144", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_bench_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.n_side),
        "bench_step": {"baoab_steps": detail["baoab_steps_per_bench_step"],
                       "note": "reference ALGORITHM restated in C/OpenMP on the host cores; a bench step is a bounded, "
                               "fully executed sample (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": detail["threads"], "kind": "port",
                         "sample": sample, **detail},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------
# replica exchange: BASELINE.json config 5 (64 temperatures x LJ N=8192), replicas sharded over the ranks
# ---------------------------------------------------------------------------------------------------
def bench_remd(dev, rank, world, sweeps=20, warmup=3, n_replicas=64, steps_per_sweep=100, cpu_baseline=False):
    import hashlib
    import torch
    import torch.distributed as dist
    from chiron_b200 import unit
    from chiron_b200.mcmc import LangevinDynamicsMove, MCMCSampler, MoveSchedule
    from chiron_b200.multistate import MultiStateSampler
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.reporters import MultistateReporter
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.testsystems import LennardJonesFluid
    from chiron_b200 import random as crandom
    # replica k starts from its own jittered lattice (seed 5 + k)
    ljs = [LennardJonesFluid(cells=(16, 16, 32), reduced_density=RHO_STAR, sigma=SIGMA * unit.nanometer,
                             epsilon=EPS_KCAL * unit.kilocalories_per_mole, seed=5 + k) for k in range(n_replicas)]
    lj = ljs[0]
    potential = LJPotential(lj.topology, lj.sigma, lj.epsilon, RC * unit.nanometer)
    temps = [TEMP_K * 2.0 ** (k / (n_replicas - 1.0)) for k in range(n_replicas)]
    thermo = [ThermodynamicState(potential, temperature=t * unit.kelvin) for t in temps]
    keys = crandom.split(crandom.PRNGKey(99), n_replicas)
    states = [SamplerState(ljs[k].positions, keys[k], box_vectors=lj.box_vectors) for k in range(n_replicas)]
    nbrs = [NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=RC * unit.nanometer, skin=SKIN * unit.nanometer,
                              n_max_neighbors=400, builder="cell") for _ in range(n_replicas)]
    move = LangevinDynamicsMove(timestep=DT_PS * unit.picosecond, collision_rate=GAMMA / unit.picosecond,
                                number_of_steps=steps_per_sweep)
    ms = MultiStateSampler(MCMCSampler(MoveSchedule([("LangevinDynamicsMove", move)])), MultistateReporter(),
                           exchange="neighbors", exchange_seed=17, mcmc_iterations_per_sweep=1)
    ms.create(thermo, states, nbrs)
    ms._offline_estimator = None      # MBAR is post-hoc analysis (pymbar in the reference), not part of a sweep
    ms._online_estimator = type("NoAnalysis", (), {"update": lambda self: None, "f_k": None})()
    ms._reporter._default_properties = ["u_kn", "state_index"]
    ms.run(warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms.reset_phase_timers()
    eng = ms._batched.engine if ms._batched else None
    run_wall = [0.0]
    if eng is not None:
        eng.step_timing(reset=True)
        st0 = eng.stats()
        _run = eng.run

        def timed_run(*a, **k):        # wall time inside chx_ljmd_run (it ends with a host read of the keys)
            tr = time.perf_counter()
            out_ = _run(*a, **k)
            run_wall[0] += time.perf_counter() - tr
            return out_
        eng.run = timed_run
    t0 = time.perf_counter()
    ms.run(warmup + sweeps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    acc, prop = int(ms._n_accepted_matrix.sum()), int(ms._n_proposed_matrix.sum())
    batched = bool(ms._batched)
    # per-phase breakdown of a sweep on this rank (host wall time; every phase ends with a host read)
    ph = ms.phase_seconds
    per = 1e3 / max(1, ph["sweeps"])
    phases = {"mix_ms": ph["mix"] * per, "propagate_ms": ph["propagate"] * per,
              "energies_and_exchange_ms": ph["energies_and_exchange"] * per,
              "report_ms": ph["report_and_analysis"] * per}
    if eng is not None:
        kms, ksteps = eng.step_timing()
        st1 = eng.stats()
        phases["step_kernels_ms"] = (kms / ksteps * steps_per_sweep) if ksteps else None
        phases["engine_groups"] = int(getattr(eng, "G", 1))   # engines on their own streams / host threads
        phases["table_builds_per_sweep"] = (st1["table_rebuilds"] - st0["table_rebuilds"]) / max(1, sweeps)
        if phases["step_kernels_ms"] is not None:
            phases["builds_and_host_in_propagate_ms"] = phases["propagate_ms"] - phases["step_kernels_ms"]
        phases["local_replicas"] = eng.R
        phases["engine_run_ms"] = run_wall[0] * per
        phases["python_in_propagate_ms"] = phases["propagate_ms"] - phases["engine_run_ms"]
    # fixed-seed fingerprint: state indices after the run (swap decisions are taken identically on every rank from a
    # shared counter key) and the energy matrix rounded to 1e-6 relative (a replica's trajectory does not depend on
    # the rank that runs it beyond fp32 summation order): compare the fingerprints of runs at different N
    u = np.asarray(ms._energy_thermodynamic_states, dtype=np.float64)
    fp_states = hashlib.sha1(np.asarray(ms._replica_thermodynamic_states, dtype=np.int64).tobytes()).hexdigest()[:12]
    out = {"sweeps_per_s": sweeps / dt, "ms_per_sweep": 1e3 * dt / sweeps, "replicas": n_replicas,
           "particles_per_replica": lj.n_particles, "baoab_steps_per_sweep": steps_per_sweep,
           "replica_steps_per_s": sweeps * n_replicas * steps_per_sweep / dt,
           "exchange": "even/odd neighbour swaps; rows of the reduced-potential matrix stay on the device, one NCCL "
                       "all_gather_into_tensor + one D2H copy of the 64x64 matrix per sweep",
           "last_sweep_swaps_accepted": acc, "last_sweep_swaps_proposed": prop, "batched_engine": batched,
           "sweeps": sweeps, "api": "MultiStateSampler.run", "phases_rank0": phases,
           "fingerprint": {"state_indices_sha1": fp_states, "state_indices": [int(v) for v in ms._replica_thermodynamic_states],
                           "u_diag_sum": float(np.trace(u)), "u_sum": float(u.sum())}}
    if cpu_baseline and rank == 0:
        # the reference algorithm (C/OpenMP port, all host threads): 2 replicas x one sweep, scaled to 64 replicas
        try:
            from oracle import cport, dynamics as dyn, jax_random as jr
            cport.use_all_cores()
            x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=np.float32)
            box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=np.float32)
            n = x.shape[0]
            v0 = dyn.maxwell_boltzmann(jr.PRNGKey(11), np.full(n, MASS), TEMP_K)
            tt = []
            for rep in range(2):
                tc = time.perf_counter()
                cport.langevin_lj(x, v0, np.full(n, MASS, np.float32), box, SIGMA, EPS, RC, SKIN, 400,
                                  8.314462618e-3 * temps[rep * (n_replicas - 1)], DT_PS, GAMMA,
                                  np.array([0, 99 + rep], np.uint32), steps_per_sweep)
                tt.append(time.perf_counter() - tc)
            t_sweep = float(np.mean(tt)) * n_replicas
            out["cpu_baseline"] = {"value": 1.0 / t_sweep, "unit": "sweeps/s", "cores": cport.num_threads(), "kind": "port",
                                   "sample": "2 replicas (coldest, hottest) x %d full BAOAB steps incl. the reference O(N^2) "
                                             "list builds at N=%d, scaled to %d replicas run one after the other (the "
                                             "reference propagates replicas serially, multistate.py:497-510); energy matrix "
                                             "and swaps not included" % (steps_per_sweep, n, n_replicas)}
        except Exception as exc:
            out["cpu_baseline"] = {"error": repr(exc)}
    return out


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # chiron logs through loguru, whose default sink prints every DEBUG / INFO record (~0.1 ms each, six per
    # replica-exchange sweep): a production run keeps warnings only
    from loguru import logger
    logger.remove()
    logger.add(sys.stderr, level="WARNING")

    from chiron_b200 import build as _build
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    from chiron_b200 import _lib, unit
    from chiron_b200._engine import LJLangevinEngine
    from chiron_b200.integrators import LangevinIntegrator
    from chiron_b200.neighbors import NeighborListNsqrd, OrthogonalPeriodicSpace
    from chiron_b200.potential import LJPotential
    from chiron_b200.states import SamplerState, ThermodynamicState
    from chiron_b200.utils import PRNG, kT_md
    import ctypes as C

    lj, x, box = make_system(args.n_side, seed=4 + rank)
    n = x.shape[0]
    kT = kT_md(TEMP_K * unit.kelvin)
    ctx = _lib.get_context(dev)
    S, K, W = args.inner, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------
    eng = LJLangevinEngine(n, np.diag(box), SIGMA, EPS, RC, SKIN, DT_PS, GAMMA, kT,
                           internal_skin=args.internal_skin, device=dev)
    from chiron_b200 import random as crandom
    from chiron_b200.utils import initialize_velocities
    v0 = initialize_velocities(TEMP_K * unit.kelvin, lj.topology, crandom.PRNGKey(11 + rank))
    v0 = v0.value_in_unit_system(unit.md_unit_system).cpu().numpy()
    mass = np.full(n, MASS, np.float32)
    eng.set_state(x, v0, mass, [kT])
    keys = crandom.PRNGKey(1234 + rank).reshape(1, 2)
    for _ in range(W):
        keys, _ = eng.run(S, keys)
    st0 = eng.stats()
    eng.step_timing(reset=True)
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(K):
            keys, _ = eng.run(S, keys)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launches - launches0
    step_kernel_ms_total, step_kernel_launches = eng.step_timing()
    st1 = eng.stats()
    e_now = float(eng.energy()[0])
    st_e = eng.stats()
    p_int, p_cand_internal = st_e["interacting_pairs"], st_e["candidate_pairs"]
    # P_cand of SURVEY.md section 8d: size of the REFERENCE Verlet list (i<j, d < rc+skin, exact
    # predicate) for the current positions, from the reference-shaped cell-list builder
    ref_list = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=RC * unit.nanometer, skin=SKIN * unit.nanometer,
                                 n_max_neighbors=400, builder="cell")
    x_now = eng.get_state()[0]
    ref_list.build(x_now, box)
    p_cand = int(ref_list.n_neighbors.sum().item())
    del ref_list
    torch.cuda.empty_cache()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    steps_per_s = world * K * S / (ms_max * 1e-3)

    # ---- dominant kernel: force kernel duration, FP32 peak ----------------------------------------
    reps = 50
    eng.force_only(5)
    torch.cuda.synchronize()
    ev0.record(); eng.force_only(reps); ev1.record(); torch.cuda.synchronize()
    force_ms = ev0.elapsed_time(ev1) / reps
    flops = C.c_double(0.0)
    ctx.call("chx_fma_peak", 2000, C.byref(flops))
    torch.cuda.synchronize()
    ev0.record(); ctx.call("chx_fma_peak", 20000, C.byref(flops)); ev1.record(); torch.cuda.synchronize()
    fp32_peak_tflops = flops.value / (ev0.elapsed_time(ev1) * 1e-3) / 1e12
    ctx.call("chx_fma_peak", -2000, C.byref(flops))
    torch.cuda.synchronize()
    ev0.record(); ctx.call("chx_fma_peak", -20000, C.byref(flops)); ev1.record(); torch.cuda.synchronize()
    fp32x2_peak_tflops = flops.value / (ev0.elapsed_time(ev1) * 1e-3) / 1e12
    pair_flops = 18.0 * p_cand + 21.0 * p_int
    step_flops = pair_flops + 140.0 * n
    # dominant kernel = the fused step kernel (pair loop + BAOAB update): duration from CUDA events around
    # the graph replays of the timed region in which every launch did its full work (chx_ljmd_step_timing)
    if step_kernel_launches:
        step_kernel_ms = step_kernel_ms_total / step_kernel_launches
    else:
        # fewer than one chunk per call (--inner < 8): no graph replay was timed; the whole-step
        # time is an upper bound of the kernel's duration
        step_kernel_ms = ms_max / (K * S)
    achieved_tflops = step_flops / (step_kernel_ms * 1e-3) / 1e12
    pair_loop_tflops = pair_flops / (force_ms * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_bytes = 4.0 * p_cand + 80.0 * n

    # ---- end to end through the public API with host buffers ---------------------------------------
    e2e = None
    if not args.no_e2e:
        potential = LJPotential(lj.topology, lj.sigma, lj.epsilon, RC * unit.nanometer)
        nbr = NeighborListNsqrd(OrthogonalPeriodicSpace(), cutoff=RC * unit.nanometer, skin=SKIN * unit.nanometer,
                                n_max_neighbors=400, builder="cell")
        ts = ThermodynamicState(potential, temperature=TEMP_K * unit.kelvin)
        integ = LangevinIntegrator(timestep=DT_PS * unit.picosecond, collision_rate=GAMMA / unit.picosecond)
        hx = torch.from_numpy(x).pin_memory()
        hv = torch.from_numpy(v0).pin_memory()
        PRNG.set_seed(1234 + rank)
        key = PRNG.get_random_key()
        # two pairs of pinned host buffers: the result of call k is the input of call k + 1
        bufs = [(hx, hv), (torch.empty((n, 3), dtype=torch.float32).pin_memory(),
                           torch.empty((n, 3), dtype=torch.float32).pin_memory())]
        turn = [0]

        def one_call(key):
            (ix, iv), (ox, ov) = bufs[turn[0]], bufs[1 - turn[0]]
            state = SamplerState(unit.Quantity(ix.to(dev, non_blocking=True), unit.nanometer), key,
                                 velocities=unit.Quantity(iv.to(dev, non_blocking=True), unit.nanometer / unit.picosecond),
                                 box_vectors=unit.Quantity(box, unit.nanometer))
            out, _ = integ.run(state, ts, number_of_steps=S, nbr_list=nbr)
            ox.copy_(out.positions, non_blocking=True)
            ov.copy_(out.velocities, non_blocking=True)
            energy = float(integ._engine.energy()[0])      # D2H read of the step's result (synchronises)
            torch.cuda.synchronize()
            turn[0] = 1 - turn[0]
            return out._current_PRNG_key, energy
        for _ in range(min(W, 2)):
            key, _ = one_call(key)
        barrier()
        t0 = time.perf_counter()
        ke = max(2, min(K, 10))
        for _ in range(ke):
            key, e_last = one_call(key)
        barrier()
        dt_e2e = time.perf_counter() - t0
        t_e = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * ke * S / float(t_e.item()), "unit": "steps/s",
               "h2d_bytes_per_step": int(2 * n * 12 + 36), "d2h_bytes_per_step": int(2 * n * 12 + 8),
               "calls": ke, "api": "LangevinIntegrator.run(SamplerState[host], ThermodynamicState, number_of_steps=%d, nbr_list)" % S}

    # ---- replica exchange (config 5), replicas sharded over the ranks -----------------------------------
    remd = None
    if not args.no_remd:
        del eng
        torch.cuda.empty_cache()
        try:
            remd = bench_remd(dev, rank, world, sweeps=args.remd_sweeps, cpu_baseline=(world == 1 and not args.no_cpu_baseline))
        except Exception as exc:   # the headline number must survive a failure of the secondary workload
            remd = {"error": repr(exc)}

    # ---- Monte Carlo configs 2 and 3 (rank 0, N = 1 only): secondary numbers ------------------------------
    mc = None
    if rank == 0 and world == 1 and not args.no_mc:
        try:
            sys.path.insert(0, os.path.join(ROOT, "profiles"))
            import bench_mc
            mc = bench_mc.main(32, quiet=True, cpu_baselines=not args.no_cpu_baseline)
        except Exception as exc:
            mc = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, bounded sample) ------------------------------------------------------
    cpu = None
    ref_events = st1["reference_rebuilds"] - st0["reference_rebuilds"]
    ref_interval = (K * S / ref_events) if ref_events > 0 else None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import __graft_entry__ as ge
        ge.build_oracle()
        # bounded sample (~15 s): 4 executed bench steps of the reference arm; the rebuild interval is the one the
        # engine's exact tracker of the reference condition measured in the timed region above
        from oracle import cport
        cport.use_all_cores()
        v, _, detail, _ = cpu_reference_bench_steps(x, box, v0, 4, 1, ref_interval)
        cpu = {"value": v, "unit": "steps/s", "cores": detail["threads"], "kind": "port",
               "sample": CPU_SAMPLE % (4, n, detail["baoab_steps_per_bench_step"], max(1, detail["build_row_stride"]),
                                       max(1, detail["build_row_stride"]),
                                       ("%.0f" % ref_interval) if ref_interval else "no event in the timed region: never"),
               **detail}

    if rank == 0:
        rebuilds = st1["table_rebuilds"] - st0["table_rebuilds"]
        line = {
            "metric": "LJ Langevin steps/s at N=262144", "value": steps_per_s, "unit": "steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args.n_side),
            "bench_step": {"baoab_steps": S,
                           "l2": "no flush between steps: the inputs of a step are produced by the previous one; tables "
                                 "(%.0f MB) + state are L2 resident inside the loop, which is the steady state of this workload"
                                 % (st_e["blocks"] * st_e["tile_capacity"] * st_e["tile_bytes"] / 1e6),
                           "internal_skin_nm": args.internal_skin or round(0.35 * SIGMA, 4)},
            "pair_interactions_per_s": p_int * steps_per_s / world * world,
            "pair_tests_per_s": p_cand * steps_per_s,
            "p_cand": p_cand, "p_int": p_int, "p_cand_internal_tables": p_cand_internal, "potential_energy_kj_mol": e_now,
            "table_rebuilds_in_timed_region": rebuilds,
            "reference_rebuild_interval_steps": ref_interval,
            "lane_utilisation": st_e.get("lane_utilisation"),
            "ms_per_baoab_step": ms_max / (K * S),
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "kernel": "k_md_force<UPDATE> (one launch per Langevin step: pair loop + BAOAB)",
                         "achieved": achieved_tflops,
                         "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak_tflops,
                         "peak_source": "FFMA-chain microbenchmark measured in this run (MEASURED_PEAKS.json has no fp32 figure)",
                         "ffma2_chain_tflops": fp32x2_peak_tflops, "flops_per_launch": step_flops,
                         "kernel_ms": step_kernel_ms, "kernel_launches_timed": int(step_kernel_launches),
                         "traffic": ncu_step_kernel_traffic(),
                         "traffic_source": "profiles/r02_step_kernel_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full --clock-control none --cache-control none (WARM L2: launch 400 of the step loop; tables and state are L2 resident there; a cold-L2 launch reads 70 MB, profiles/r01_step_kernel_ncu.md)",
                         "pair_loop_only": {"kernel": "k_md_force (no update)", "kernel_ms": force_ms, "flops_per_launch": pair_flops,
                                            "achieved": pair_loop_tflops, "frac": pair_loop_tflops / fp32_peak_tflops},
                         "step_flops": step_flops,
                         "step_frac": step_flops / (ms_max / (K * S) * 1e-3) / 1e12 / fp32_peak_tflops,
                         "hbm": {"achieved": hbm_bytes / (ms_max / (K * S) * 1e-3) / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "bytes_per_step": hbm_bytes,
                                 "frac": hbm_bytes / (ms_max / (K * S) * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}},
            "cpu_baseline": cpu,
            "remd": remd,
            "mc": mc,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything libraries print while the bench
    runs (NCCL's version banner, loguru, nvcc) was redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    a = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)           # C libraries included
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
