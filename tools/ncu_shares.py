"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the profiled window."""
import collections
import csv
import re
import sys


def main(path):
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines[start:]):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:72]
        if "spin_kernel" in name:   # bench.py's device-side sleep in front of each instrumented step: not part of the step
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms (cold-cache, serialised: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  avg {1e3 * v[1] / v[0]:9.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
