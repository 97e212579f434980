"""Profiling driver: the bench workload (LJ argon N=262,144, rho*=0.8), a few warm steps, then
`repeats` launches of the force kernel on the current positions (every one of them does real work).

    ncu --set full --clock-control none --import-source on -k regex:k_md_force -s 3 -c 1 \
        -o gpurun_out/force python profiles/prof_force.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(n_side=64, steps=0, repeats=6):
    import torch
    from chiron_b200 import random as crandom, unit
    from chiron_b200._engine import LJLangevinEngine
    from chiron_b200.utils import initialize_velocities, kT_md
    dev = torch.device("cuda", 0)
    lj, x, box = bench.make_system(n_side, seed=4)
    n = x.shape[0]
    kT = kT_md(bench.TEMP_K * unit.kelvin)
    skin_int = float(os.environ["INTERNAL_SKIN"]) if "INTERNAL_SKIN" in os.environ else None
    eng = LJLangevinEngine(n, np.diag(box), bench.SIGMA, bench.EPS, bench.RC, bench.SKIN, bench.DT_PS,
                           bench.GAMMA, kT, internal_skin=skin_int, device=dev)
    v0 = initialize_velocities(bench.TEMP_K * unit.kelvin, lj.topology, crandom.PRNGKey(11))
    v0 = v0.value_in_unit_system(unit.md_unit_system).cpu().numpy()
    eng.set_state(x, v0, np.full(n, bench.MASS, np.float32), [kT])
    keys = crandom.PRNGKey(1234).reshape(1, 2)
    if steps:
        keys, _ = eng.run(steps, keys)
    eng.force_only(repeats)
    torch.cuda.synchronize()
    if os.environ.get("TIME", "0") == "1":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.force_only(50); e1.record(); torch.cuda.synchronize()
        t_force = e0.elapsed_time(e1) / 50 * 1e3
        keys, _ = eng.run(200, keys)
        torch.cuda.synchronize()
        e0.record(); keys, _ = eng.run(1000, keys); e1.record(); torch.cuda.synchronize()
        t_step = e0.elapsed_time(e1)
        print("TIMING tag=%s force_us=%.2f step_us=%.2f steps_per_s=%.0f" % (
            os.environ.get("TAG", ""), t_force, t_step, 1e6 / t_step))
    print(eng.stats())


if __name__ == "__main__":
    main(int(os.environ.get("N_SIDE", "64")), int(os.environ.get("STEPS", "0")), int(os.environ.get("REPEATS", "6")))
