"""Occupancy over time of ONE step-kernel launch (library built with -DCHX_TRACE=1): every warp records its start /
end globaltimer and SM.  Prints the span of the launch, how long the machine takes to fill and to drain, and the
block durations of the first and the later CTAs."""
import ctypes as C
import os
import sys

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
    cells = tuple(int(c) for c in os.environ.get("CELLS", "64,64,64").split(","))
    R = int(os.environ.get("NREP", "1"))
    lj = LennardJonesFluid(cells=cells, reduced_density=bench.RHO_STAR, sigma=bench.SIGMA * unit.nanometer,
                           epsilon=bench.EPS_KCAL * unit.kilocalories_per_mole, seed=5)
    x = np.asarray(lj.positions.value_in_unit(unit.nanometer), dtype=np.float32)
    box = np.asarray(lj.box_vectors.value_in_unit(unit.nanometer), dtype=np.float32)
    n = x.shape[0]
    kTs = [kT_md(bench.TEMP_K * 2.0 ** (k / max(R - 1.0, 1.0)) * unit.kelvin) for k in range(R)]
    eng = LJLangevinEngine(n, np.diag(box), bench.SIGMA, bench.EPS, bench.RC, bench.SKIN, bench.DT_PS,
                           bench.GAMMA, kTs[0], n_replicas=R, device=dev)
    v0 = initialize_velocities(bench.TEMP_K * unit.kelvin, lj.topology, crandom.PRNGKey(11))
    v0 = v0.value_in_unit_system(unit.md_unit_system).cpu().numpy()
    eng.set_state(np.tile(x[None], (R, 1, 1)), np.tile(v0[None], (R, 1, 1)), np.full(n, bench.MASS, np.float32), kTs)
    keys = np.asarray(crandom.split(crandom.PRNGKey(1234), R), dtype=np.uint32).reshape(R, 2)
    keys, _ = eng.run(300, keys)
    st = eng.stats()
    nblk = st["blocks"] * R
    lib = _lib.load_library()
    buf = torch.zeros((nblk * 4, 3), dtype=torch.int64, device=dev)     # room for 4 warps per block
    assert lib.chx_debug_set_trace(C.c_void_p(buf.data_ptr()), C.c_int(5)) == 0
    keys, _ = eng.run(8, keys)          # steps 0..7 of this run: one graph chunk; step 5 is traced
    torch.cuda.synchronize()
    t = buf.cpu().numpy()
    t = t[t[:, 1] > 0]
    split = len(t) // nblk if len(t) % nblk == 0 and len(t) // nblk in (1, 2, 4) else 1   # cut pieces: CTAs != blocks
    lib.chx_debug_set_trace(C.c_void_p(0), C.c_int(-1))
    t0, t1, sm = t[:, 0].astype(np.float64), t[:, 1].astype(np.float64), t[:, 2]
    base = t0.min()
    t0, t1 = (t0 - base) / 1e3, (t1 - base) / 1e3      # us
    span = t1.max()
    print("TRACE warps=%d split=%d span_us=%.2f" % (len(t0), split, span))
    # active warps over time
    grid = np.linspace(0.0, span, 201)
    act = np.array([((t0 <= g) & (t1 > g)).sum() for g in grid])
    peak = act.max()
    fill90 = grid[np.argmax(act >= 0.9 * peak)]
    last90 = grid[len(act) - 1 - np.argmax(act[::-1] >= 0.9 * peak)]
    last50 = grid[len(act) - 1 - np.argmax(act[::-1] >= 0.5 * peak)]
    print("TRACE peak_active=%d reaches_90pct_at=%.2f us; falls_below_90pct_at=%.2f, below_50pct_at=%.2f, end=%.2f" % (
        peak, fill90, last90, last50, span))
    print("TRACE active warps at 0,5,..100%% of the span: %s" % " ".join(str(int(a)) for a in act[::10]))
    dur = t1 - t0
    order = np.argsort(t0)
    first = order[:min(len(order), 4736 // 1)]
    later = order[len(first):]
    print("TRACE start times: first warp 0.00, median %.2f, 90th pct %.2f, last %.2f us" % (
        np.median(t0), np.percentile(t0, 90), t0.max()))
    print("TRACE durations us: all median %.2f (min %.2f max %.2f); first %d started: median %.2f; later %d: median %.2f" % (
        np.median(dur), dur.min(), dur.max(), len(first), np.median(dur[first]), len(later),
        np.median(dur[later]) if len(later) else float("nan")))
    # integral of active warps = warp-us of work; work / span = average occupancy
    print("TRACE warp_us=%.0f avg_active=%.0f  (work / peak = %.2f us of a full machine)" % (
        dur.sum(), dur.sum() / span, dur.sum() / peak))
    # launch position -> duration (deciles of the launch order), and the work estimate's predictive power
    nd = 10
    cta_dur = dur.reshape(-1, split).max(axis=1) if split > 1 else dur
    dec = [float(np.mean(c)) for c in np.array_split(cta_dur, nd)]
    print("TRACE mean CTA duration by launch-order decile: %s" % " ".join("%.1f" % d for d in dec))
    print("TRACE CTA duration percentiles 1/10/50/90/99: %s" % " ".join("%.1f" % np.percentile(cta_dur, q) for q in (1, 10, 50, 90, 99)))
    # per SM: end of its last warp
    ends = np.array([t1[sm == k].max() for k in np.unique(sm)])
    print("TRACE per-SM last end: min %.2f median %.2f max %.2f us (%d SMs)" % (ends.min(), np.median(ends), ends.max(), len(ends)))


main()
