"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python profiles/summarize_launches.py gpurun_out/launches.csv"""
import collections
import csv
import statistics
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        agg[r[ki].split("(")[0]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':44s} {'n':>5s} {'sum us':>10s} {'share':>6s} {'median':>8s} {'max':>8s} {'min':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:44]:44s} {len(v):5d} {sum(v):10.1f} {sum(v) / tot:6.1%} {statistics.median(v):8.1f} {max(v):8.1f} {min(v):7.1f}")
    print(f"total {tot:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1])
