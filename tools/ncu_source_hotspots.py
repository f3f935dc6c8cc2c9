"""Aggregate warp-stall samples per CUDA source line from an ncu report.

usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
       python tools/ncu_source_hotspots.py src.csv [top]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    fname = "?"
    hdr = None
    per_line = defaultdict(lambda: defaultdict(float))
    src_text = {}
    total = 0.0
    for r in csv.reader(open(path, errors="replace")):
        if not r:
            continue
        if r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        key = (fname, line)
        src_text.setdefault(key, r[1].strip())
        try:
            samples = float(r[hdr.index("# Samples")] or 0)
        except ValueError:
            samples = 0.0
        per_line[key]["samples"] += samples
        total += samples
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    per_line[key][h] += float(r[i] or 0)
                except ValueError:
                    pass
    print(f"total samples {total:.0f}")
    for key, d in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        st = sorted(((v, k) for k, v in d.items() if k != "samples" and v > 0), reverse=True)[:3]
        sts = " ".join(f"{k[6:]}={100 * v / max(d['samples'], 1):.0f}%" for v, k in st)
        print(f"{100 * d['samples'] / total:5.1f}%  {key[0]}:{key[1]:<5d} {sts:<44s} | {src_text[key][:90]}")


if __name__ == "__main__":
    main()
