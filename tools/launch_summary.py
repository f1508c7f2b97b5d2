"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time and share."""
import collections
import csv
import sys


def main(path, steps):
    rows = []
    with open(path) as f:
        rd = csv.reader(l for l in f if not l.startswith("=="))
        hdr = next(rd)
        for r in rd:
            if len(r) == len(hdr):
                rows.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = r["Kernel Name"][:96]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1000 if u == "us" else (v / 1e6 if u == "ns" else v)
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches, {tot:.3f} ms total under ncu (cold-cache, serialised: compare shares), ~{steps} steps")
    print(f"# {'ms/step':>9} {'launches':>8} {'share':>6}  kernel")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t / steps:10.4f} {c:8d} {100 * t / tot:5.1f}%  {n}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
