"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
Usage: python tools/launch_summary.py launches.csv [steps_in_capture]"""
import csv
import re
import sys
from collections import OrderedDict


def main(path: str, steps: float) -> None:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    agg: "OrderedDict[str, list]" = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        name = re.sub(r"<unnamed>::", "", name)[:110]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e3
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"Kernel time per step: **{tot / steps:.0f} us** over {n / steps:.0f} launches ({len(rows)} launches captured = {steps:g} steps)\n")
    print("| kernel | launches / step | us / step | share |\n|---|---|---|---|")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c / steps:g} | {us / steps:.1f} | {100 * us / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
