"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of GPU time)."""
import collections
import csv
import io
import sys


def main(path, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[row["Kernel Name"]][0] += 1
        agg[row["Kernel Name"]][1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    n = sum(v[0] for v in agg.values())
    print(f"{n} launches, {tot / 1e6:.3f} ms total GPU time (cold-cache, serialised: compare shares)")
    ours = sum(v[1] for k, v in agg.items() if "gemnet::" in k)
    print(f"gemnet:: kernels: {ours / tot * 100:.1f}% of GPU time, {sum(v[0] for k, v in agg.items() if 'gemnet::' in k)} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1] / tot * 100:6.2f}%  n={v[0]:5d}  avg={v[1] / v[0] / 1e3:9.2f}us  {k[:120]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
