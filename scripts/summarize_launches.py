"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into the markdown table kept under profiles/.
usage: python scripts/summarize_launches.py launches.csv [launches_per_forward]"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("ddpm::", "").replace("(bool)", "")


def main():
    path = sys.argv[1]
    per_fwd = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        grid = r.get("Grid Size", "").replace(" ", "")
        rows.append((short(r["Kernel Name"]) + f" grid={grid}", us))
    total = sum(u for _, u in rows)
    agg = OrderedDict()
    for k, u in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += u
    print(f"{len(rows)} launches, {total:.0f} us serialised\n")
    print("| share | launches | avg us | kernel |\n|---|---|---|---|")
    for k, (n, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * u / total:.2f}% | {n} | {u / n:.1f} | `{k}` |")
    if per_fwd:
        # first whole forward: starts at the first conv_in launch
        start = next(i for i, (k, _) in enumerate(rows) if k.startswith("conv_in"))
        fwd = rows[start:start + per_fwd]
        print(f"\nOne forward ({per_fwd} launches, {sum(u for _, u in fwd):.0f} us serialised), in launch order:\n\n```")
        for k, u in fwd:
            print(f"{u:8.1f}  {k}")
        print("```")


if __name__ == "__main__":
    main()
