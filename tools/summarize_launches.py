"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

usage: python tools/summarize_launches.py launches.csv "header line" > summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    header = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        name = r["Kernel Name"]
        name = re.sub(r"^void\s+", "", name)
        name = re.sub(r"\(.*$", "", name)
        name = re.sub(r"^mgbx::", "", name)
        name = re.sub(r"cub::(CUB_\w+::)?", "cub::", name)
        name = re.sub(r"(cub::\w+)<.*", r"\1<...>", name)
        rows.append((name, us))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    print("# " + header)
    print(f"{'kernel':<42}{'launches':>9}{'total us':>13}{'share':>8}{'avg us':>11}")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:41]:<42}{c:>9}{us:>13.1f}{100 * us / tot:>7.1f}%{us / c:>11.1f}")
    print(f"{'total':<42}{len(rows):>9}{tot:>13.1f}")


if __name__ == "__main__":
    main()
