"""Debug: persistent vs per-launch step loop on the small-skin system of test_fused_engine_rebuild_events."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from chiron_b200 import random as crandom
from chiron_b200._engine import LJLangevinEngine
import test_gpu_parity as T

lj_sys, x, box = T._lj_system(8, 0.8, seed=43)
n = x.shape[0]
rng = np.random.default_rng(5)
v = rng.normal(0, 0.25, (n, 3)).astype(np.float32)
mass = np.full(n, 39.948, np.float32)
res = {}
for mode in ("launch", "persist", "persist_olddeal"):
    os.environ["CHX_MD_PERSIST"] = "0" if mode == "launch" else "1"
    os.environ["CHX_MD_OLD_DEAL"] = "1" if mode == "persist_olddeal" else "0"
    out = []
    for nsteps in (1, 2, 3, 5, 8, 12, 20, 30, 60):
        eng = LJLangevinEngine(n, np.diag(box), 0.34, 0.238 * 4.184, 1.02, 0.02, 0.002, 1.0, 2.494, device="cuda:0")
        eng.set_state(x, v, mass, [2.494])
        keys, _ = eng.run(nsteps, crandom.PRNGKey(5).reshape(1, 2))
        xs, vs, f, ref = eng.get_state(want_force=True, want_ref=True)
        st = eng.stats()
        out.append((nsteps, xs.cpu().numpy(), vs.cpu().numpy(), f.cpu().numpy(), st["table_rebuilds"], st["reference_rebuilds"], keys.copy()))
        eng.close()
    res[mode] = out
for mode in ("persist", "persist_olddeal"):
    for a, b in zip(res["launch"], res[mode]):
        dx = a[1] - b[1]; L = np.diag(box); dx -= L * np.round(dx / L)
        print(mode, "n=%d dx=%.3g dv=%.3g df=%.3g tr=%d/%d rr=%d/%d key=%s" % (a[0], np.abs(dx).max(), np.abs(a[2] - b[2]).max(),
              np.abs(a[3] - b[3]).max() / np.abs(a[3]).max(), a[4], b[4], a[5], b[5], np.array_equal(a[6], b[6])))
