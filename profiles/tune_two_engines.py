"""8 replicas x 8,192 particles on ONE GPU as G engines of 8/G replicas, each on its own stream and host thread
(ctypes releases the GIL): do the latency-bound phases of one engine (table rebuild, kernel ramp and tail) fill
with the work of the other?  Prints us per step of the whole set (100-step runs, like a REMD sweep)."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from chiron_b200 import _lib, random as crandom, unit
    from chiron_b200._engine import LJLangevinEngine
    from chiron_b200.testsystems import LennardJonesFluid
    from chiron_b200.utils import initialize_velocities, kT_md
    dev = torch.device("cuda", 0)
    R = int(os.environ.get("NREP", "8"))
    G = int(os.environ.get("NGROUPS", "2"))
    sweeps = int(os.environ.get("SWEEPS", "20"))
    lj = LennardJonesFluid(cells=(16, 16, 32), reduced_density=bench.RHO_STAR, sigma=bench.SIGMA * unit.nanometer,
                           epsilon=bench.EPS_KCAL * unit.kilocalories_per_mole, seed=5)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=np.float32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=np.float32)
    n = x.shape[0]
    temps = [bench.TEMP_K * 2.0 ** (k / max(R - 1.0, 1.0)) for k in range(R)]
    kTs = [kT_md(t * unit.kelvin) for t in temps]
    v0 = initialize_velocities(bench.TEMP_K * unit.kelvin, lj.topology, crandom.PRNGKey(11))
    v0 = v0.value_in_unit_system(unit.md_unit_system).cpu().numpy()
    keys = np.asarray(crandom.split(crandom.PRNGKey(1234), R), dtype=np.uint32).reshape(R, 2)
    if os.environ.get("LOCKSTEP", "0") == "1":
        # the product path: _EngineGroups.run joins all engines after every 100-step run (one REMD sweep)
        from chiron_b200.multistate import _EngineGroups
        eg = _EngineGroups(G, n, np.diag(box), bench.SIGMA, bench.EPS, bench.RC, bench.SKIN, bench.DT_PS, bench.GAMMA,
                           kTs[0], R, dev)
        if os.environ.get("NO_PHASE", "0") == "1":
            for e in eg.engines:
                e.set_chunk_phase(0, 1)
        xt = torch.as_tensor(np.tile(x[None], (R, 1, 1)), device=dev)
        vt = torch.as_tensor(np.tile(v0[None], (R, 1, 1)), device=dev)
        eg.set_state(xt, vt, torch.full((n,), bench.MASS, dtype=torch.float32, device=dev), kTs)
        for _ in range(3):
            keys, _e = eg.run(100, keys)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(sweeps):
                keys, _e = eg.run(100, keys)
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / (sweeps * 100) * 1e6)
        xs_ = eg.get_state()[0].cpu().numpy()
        print("TWO lockstep nrep=%d groups=%d phase=%s us_per_step_all=%.2f ms_per_100_steps=%.3f checksum=%.6f" % (
            R, G, os.environ.get("NO_PHASE", "0") != "1", best, best / 10.0, float(np.abs(xs_).sum())))
        return
    groups = [list(range(g, R, G)) for g in range(G)]        # strided, like the rank sharding
    engines, streams, gkeys = [], [], []
    for ids in groups:
        s = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(s):
            ctx = _lib.Context(0) if G > 1 else None
            eng = LJLangevinEngine(n, np.diag(box), bench.SIGMA, bench.EPS, bench.RC, bench.SKIN, bench.DT_PS,
                                   bench.GAMMA, kTs[0], n_replicas=len(ids), device=dev, ctx=ctx)
            eng.set_state(np.tile(x[None], (len(ids), 1, 1)), np.tile(v0[None], (len(ids), 1, 1)),
                          np.full(n, bench.MASS, np.float32), [kTs[i] for i in ids])
        engines.append(eng); streams.append(s); gkeys.append(keys[ids].copy())
    torch.cuda.synchronize()

    def work(g, nsw):
        with torch.cuda.stream(streams[g]):
            for _ in range(nsw):
                gkeys[g], _e = engines[g].run(100, gkeys[g])
            streams[g].synchronize()

    def sweep_all(nsw):
        if G == 1:
            work(0, nsw)
            return
        th = [threading.Thread(target=work, args=(g, nsw)) for g in range(G)]
        for t in th: t.start()
        for t in th: t.join()

    sweep_all(3)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        sweep_all(sweeps)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / (sweeps * 100) * 1e6)
    x_all = np.concatenate([e.get_state()[0].reshape(-1, n, 3).cpu().numpy() for e in engines])
    print("TWO nrep=%d groups=%d split=%s us_per_step_all=%.2f ms_per_100_steps=%.3f checksum=%.6f" % (
        R, G, os.environ.get("CHX_FORCE_SPLIT", "-"), best, best / 10.0, float(np.abs(x_all).sum())))


main()
