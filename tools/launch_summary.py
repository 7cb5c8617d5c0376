"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/launch_summary.py gpurun_out/prof_r02/bench_launches.csv [--step-marker count_star_kernel --step 1]

With --step-marker the launches of ONE step are selected: from the k-th launch of the marker kernel
(the first kernel of the mesh build) up to the next one."""
import argparse
import csv
import re
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)          # drop the argument list
    name = name.replace("void ", "")
    return name[:66]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--step-marker", default=None)
    ap.add_argument("--step", type=int, default=1)
    args = ap.parse_args()
    rows = []
    with open(args.csv, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"]))
    if args.step_marker:
        marks = [k for k, r in enumerate(rows) if args.step_marker in r[1]]
        lo = marks[args.step]
        hi = marks[args.step + 1] if args.step + 1 < len(marks) else len(rows)
        rows = rows[lo:hi]
    agg = OrderedDict()
    for _, name, ns, _grid in rows:
        a = agg.setdefault(short(name), [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"{'kernel':66s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
    for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:66s} {cnt:8d} {ns * 1e-6:10.3f} {100 * ns / total:6.1f}%")
    print(f"{'total':66s} {sum(a[0] for a in agg.values()):8d} {total * 1e-6:10.3f}")


if __name__ == "__main__":
    main()
