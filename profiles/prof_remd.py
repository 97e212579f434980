"""cProfile of MultiStateSampler sweeps (config 5) to see the host-side cost per sweep."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    from loguru import logger
    logger.remove()
    dev = torch.device("cuda", 0)
    n_rep = int(os.environ.get("NREP", "64"))
    pr = cProfile.Profile()
    orig_run = None
    # reuse bench_remd but profile only the timed part: monkeypatch time.perf_counter boundaries
    t0 = time.perf_counter()
    pr.enable()
    out = bench.bench_remd(dev, 0, 1, sweeps=int(os.environ.get("SWEEPS", "20")), warmup=3, n_replicas=n_rep)
    pr.disable()
    print(out["ms_per_sweep"], "ms/sweep", time.perf_counter() - t0, "s total")
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()
