#!/usr/bin/env python
"""Benchmark of the particle hot path: LJ argon fluid, N = 262,144, rho* = 0.8, Langevin BAOAB with
the cell-list neighbour build (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--inner S] [--impl ours|reference]

A bench "step" is one `LangevinIntegrator.run`-sized chunk of S = --inner BAOAB steps (default 100);
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
WORKLOAD = "LJ argon fluid N=262144 rho*=0.8 rc=3sigma skin=0.5nm T=300K dt=1fs Langevin BAOAB, cell-list build"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--inner", type=int, default=100, help="BAOAB steps per bench step")
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n-side", type=int, default=N_SIDE)
    p.add_argument("--internal-skin", type=float, default=None)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
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
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

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
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline: the oracle on a bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------
def cpu_reference_steps_per_s(x, box, rows=1536, repeats=2):
    """Reference algorithm restated in NumPy (JAX unavailable): per-step cost = NeighborListNsqrd
    calculate over the padded (N, M) list + masked LJ energy/force + BAOAB update; the O(N^2) build is
    amortised over the measured rebuild interval.  Timed on `rows` rows of the N=262,144 system and
    scaled by N/rows (both parts are row-separable)."""
    import torch
    from oracle import pairs, potentials as pot, jax_random as jr
    n = x.shape[0]
    torch.set_num_threads(os.cpu_count() or 1)
    sel = np.arange(rows)
    t0 = time.perf_counter()
    nbr_rows = pairs.neighbor_rows(x, box, RC + SKIN, rows=sel, chunk=256)     # O(rows * N) slice of build
    t_build_rows = time.perf_counter() - t0
    M = max(r.size for r in nbr_rows) + 10
    nl = np.zeros((rows, M), dtype=np.int64)
    mask = np.zeros((rows, M), dtype=np.int32)
    for i, r in enumerate(nbr_rows):
        nl[i, :r.size] = r
        nl[i, r.size:] = r[0] if r.size else 0
        mask[i, :r.size] = 1
    best = 1e30
    for _ in range(repeats):
        t0 = time.perf_counter()
        r, d = pairs.displacement(x[sel][:, None, :], x[nl], box)
        m = (d < np.float32(RC)) & (mask != 0)
        f = np.where(m, pot.lj_pair_force_scalar(d, SIGMA, EPS), 0).astype(np.float32)
        fv = f[..., None] * r
        F = np.zeros((n, 3), dtype=np.float32)
        F[sel] += fv.sum(axis=1)
        np.subtract.at(F, nl.reshape(-1), fv.reshape(-1, 3))
        # BAOAB for the same rows (noise on the reference's stream)
        xi = jr.normal(jr.PRNGKey(1), (rows, 3))
        v = (0.5 * xi).astype(np.float32)
        v = v + np.float32(0.0005) * F[sel] / np.float32(MASS)
        xs = x[sel] + np.float32(0.0005) * v
        xs = pairs.wrap(xs, box)
        pairs.displacement(xs, x[sel], box)
        best = min(best, time.perf_counter() - t0)
    t_step = best * n / rows
    t_build = t_build_rows * n / rows
    rebuild_interval = 150.0            # measured by the GPU run at this state point (skin 0.5 nm)
    per_step = t_step + t_build / rebuild_interval
    return 1.0 / per_step, {"t_step_s": t_step, "t_build_s": t_build, "rows": rows,
                            "rebuild_interval_steps": rebuild_interval}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lj, x, box = make_system(args.n_side, seed=4)
    cores = os.cpu_count() or 1
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        v, detail = cpu_reference_steps_per_s(x, box, rows=1024, repeats=1)
        vals.append(v)
    value = float(np.median(vals))
    line = {
        "impl": "reference", "metric": "LJ Langevin steps/s at N=262144", "value": value, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.inner / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "baoab_steps_per_bench_step": args.inner,
                   "note": "reference algorithm restated in NumPy on the host (JAX/OpenMM cannot be installed offline); "
                           "bounded sample: 1024 of 262144 list rows per step, scaled by N/rows"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": "1024 rows of the padded (N,M) list per step + amortised O(N^2) build slice", **detail},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
    achieved_tflops = pair_flops / (force_ms * 1e-3) / 1e12
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
        out_x = torch.empty((n, 3), dtype=torch.float32).pin_memory()
        out_v = torch.empty((n, 3), dtype=torch.float32).pin_memory()

        def one_call(key):
            state = SamplerState(unit.Quantity(hx.to(dev, non_blocking=True), unit.nanometer), key,
                                 velocities=unit.Quantity(hv.to(dev, non_blocking=True), unit.nanometer / unit.picosecond),
                                 box_vectors=unit.Quantity(box, unit.nanometer))
            out, _ = integ.run(state, ts, number_of_steps=S, nbr_list=nbr)
            out_x.copy_(out.positions, non_blocking=True)
            out_v.copy_(out.velocities, non_blocking=True)
            energy = float(integ._engine.energy()[0])      # D2H read of the step's result
            torch.cuda.synchronize()
            hx.copy_(out_x); hv.copy_(out_v)
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

    # ---- CPU baseline (rank 0, bounded sample) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, detail = cpu_reference_steps_per_s(x, box)
        cpu = {"value": v, "unit": "steps/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "%d of %d rows of the padded (N,M) list per step (calculate + masked LJ force + BAOAB) "
                         "+ amortised O(N^2) build slice, scaled by N/rows; NumPy restatement of the reference "
                         "algorithm, JAX unavailable" % (detail["rows"], n), **detail}

    if rank == 0:
        rebuilds = st1["table_rebuilds"] - st0["table_rebuilds"]
        line = {
            "metric": "LJ Langevin steps/s at N=262144", "value": steps_per_s, "unit": "steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.n_side == N_SIDE else WORKLOAD.replace("262144", str(n)),
                       "baoab_steps_per_bench_step": S, "n_particles": n,
                       "parallelism": "1 independent system per GPU (replicas only)",
                       "l2": "working set (tiles %.0f MB + state) exceeds nothing to flush: inputs are produced by the previous step"
                             % (st_e["blocks"] * st_e["tile_capacity"] * 256 / 1e6),
                       "internal_skin_nm": args.internal_skin or round(0.35 * SIGMA, 4)},
            "pair_interactions_per_s": p_int * steps_per_s / world * world,
            "pair_tests_per_s": p_cand * steps_per_s,
            "p_cand": p_cand, "p_int": p_int, "p_cand_internal_tables": p_cand_internal, "potential_energy_kj_mol": e_now,
            "table_rebuilds_in_timed_region": rebuilds,
            "ms_per_baoab_step": ms_max / (K * S),
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "kernel": "k_md_force", "achieved": achieved_tflops,
                         "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak_tflops,
                         "peak_source": "FFMA-chain microbenchmark measured in this run (MEASURED_PEAKS.json has no fp32 figure)",
                         "ffma2_chain_tflops": fp32x2_peak_tflops, "flops_per_launch": pair_flops, "kernel_ms": force_ms, "traffic": None,
                         "step_flops": step_flops,
                         "step_frac": step_flops / (ms_max / (K * S) * 1e-3) / 1e12 / fp32_peak_tflops,
                         "hbm": {"achieved": hbm_bytes / (ms_max / (K * S) * 1e-3) / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "bytes_per_step": hbm_bytes,
                                 "frac": hbm_bytes / (ms_max / (K * S) * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s"}},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
