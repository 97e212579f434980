"""Time the force kernel and the step loop of the fused engine for one workload:

    NREP=1 CELLS=64,64,64   single system, N = 262,144 (the headline workload)
    NREP=8 CELLS=16,16,32   8 replicas x N = 8,192 (one rank's share of config 5 at 8 GPUs)
    NREP=64 CELLS=16,16,32  config 5 on one GPU

Knobs come from the environment (CHX_FORCE_SPLIT, CHX_MD_CHUNK, CHX_MD_SKIN, ...); prints one line.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from chiron_b200 import random as crandom, unit
    from chiron_b200._engine import LJLangevinEngine
    from chiron_b200.testsystems import LennardJonesFluid
    from chiron_b200.utils import initialize_velocities, kT_md
    dev = torch.device("cuda", 0)
    cells = tuple(int(c) for c in os.environ.get("CELLS", "64,64,64").split(","))
    R = int(os.environ.get("NREP", "1"))
    steps = int(os.environ.get("STEPS", "1000"))
    lj = LennardJonesFluid(cells=cells, reduced_density=bench.RHO_STAR, sigma=bench.SIGMA * unit.nanometer,
                           epsilon=bench.EPS_KCAL * unit.kilocalories_per_mole, seed=5)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=np.float32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=np.float32)
    n = x.shape[0]
    temps = [bench.TEMP_K * 2.0 ** (k / max(R - 1.0, 1.0)) for k in range(R)]
    kTs = [kT_md(t * unit.kelvin) for t in temps]
    eng = LJLangevinEngine(n, np.diag(box), bench.SIGMA, bench.EPS, bench.RC, bench.SKIN, bench.DT_PS,
                           bench.GAMMA, kTs[0], n_replicas=R, device=dev)
    v0 = initialize_velocities(bench.TEMP_K * unit.kelvin, lj.topology, crandom.PRNGKey(11))
    v0 = v0.value_in_unit_system(unit.md_unit_system).cpu().numpy()
    eng.set_state(np.tile(x[None], (R, 1, 1)), np.tile(v0[None], (R, 1, 1)), np.full(n, bench.MASS, np.float32), kTs)
    keys = np.asarray(crandom.split(crandom.PRNGKey(1234), R), dtype=np.uint32).reshape(R, 2)
    keys, _ = eng.run(300, keys)
    eng.force_only(10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.force_only(100); e1.record(); torch.cuda.synchronize()
    t_force = e0.elapsed_time(e1) / 100 * 1e3
    best = 1e30
    eng.step_timing(reset=True)
    for _ in range(3):
        e0.record(); keys, _ = eng.run(steps, keys); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps * 1e3)
    st = eng.stats()
    kms, ksteps = eng.step_timing()
    print("TUNE nrep=%d n=%d persist=%s chunk=%s skin=%s force_us=%.2f step_us=%.2f kernel_us_per_step=%.2f lane_util=%.3f rebuilds=%d cand_per_particle=%.1f trips_per_block=%.1f" % (
        R, n, os.environ.get("CHX_MD_PERSIST", "1"), os.environ.get("CHX_MD_CHUNK", "-"),
        os.environ.get("CHX_MD_SKIN", "-"), t_force, best, kms / max(ksteps, 1) * 1e3, st["lane_utilisation"],
        st["table_rebuilds"], 2.0 * st["candidate_pairs"] / (R * n), st["trip_slots"] / 64.0 / (R * st["blocks"])))


if __name__ == "__main__":
    main()
