"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name.

    python tools/agg_launches.py gpurun_out/launches.csv[.gz] [top_n]
"""
import collections
import csv
import gzip
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(l for l in fh if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = r[ki].split("(")[0][:64]
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{n:66s} {c:5d} {t / 1e6:9.3f} ms {100 * t / tot:5.1f}% avg {t / c / 1e3:8.1f} us")


if __name__ == "__main__":
    main()
