"""`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a small system: one JSON line on stdout
with the contract's keys, the SAME `config` dict as our arm prints, and a `ms_per_step` that is a measured duration
(steps x ms_per_step fits inside the wall clock of the run -- the round-1 arm reported a modelled number)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_is_consistent():
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                        "--warmup", "1", "--n-side", "12"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    wall = time.perf_counter() - t0
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(12)                      # identical in both arms
    assert d["steps"] * d["ms_per_step"] * 1e-3 <= wall               # executed, not extrapolated
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["extrapolated"] is False and cb["cores"] >= 1
    assert cb["value"] == d["value"] == d["e2e"]["value"] > 0
    assert cb["baoab_steps_per_bench_step"] == bench.REF_STEPS_PER_BENCH_STEP
    # value = BAOAB steps executed per second of bench-step time
    assert abs(d["value"] - cb["baoab_steps_per_bench_step"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
