"""Dev tool: aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python scripts/summarize_launches.py profiles/<file>.csv [> profiles/<file>.summary.txt]
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot, n = 0.0, 0
    for row in r:
        v = float(row[iv].replace(",", ""))
        v = v / 1e3 if row[iu] == "ns" else (v * 1e3 if row[iu] == "ms" else v)
        name = re.sub(r"\(.*", "", row[ik]).replace("void ", "").replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        n += 1
    print(f"# {path}: {n} launches, {tot / 1e3:.2f} ms of device time (ncu: cold-cache, serialised — compare shares)")
    print(f"{'ms':>10} {'share':>7} {'launches':>9} {'avg us':>10}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t / 1e3:10.2f} {100 * t / tot:6.2f}% {c:9d} {t / c:10.1f}  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
