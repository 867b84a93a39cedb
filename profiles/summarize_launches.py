"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean ns, share."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, row.get("Grid Size"), row.get("Block Size")])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    own = sum(a[1] for k, a in agg.items() if "rcn::" in k)
    print(f"{'n':>5} {'mean_ns':>10} {'share':>6} {'of rcn':>7}  grid / block / kernel   (rcn:: = this repo's kernels; the rest is torch set-up "
          "and the DGEMM / IGEMM peak probes of bench.py)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        mine = f"{a[1] / own:7.3f}" if "rcn::" in k else "      -"
        print(f"{a[0]:5d} {a[1] / a[0]:10.1f} {a[1] / tot:6.3f} {mine}  {a[2]} {a[3]} {k}")


if __name__ == "__main__":
    main(sys.argv[1])
