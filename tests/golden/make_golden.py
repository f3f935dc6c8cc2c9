#!/usr/bin/env python
"""Extract the reference's own golden vectors into tests/golden/reference_goldens.json.

Run in the build container only (reads /root/reference, which does not travel to the GPU box):
    python tests/golden/make_golden.py
Sources: test/runtests.jl:13-52 and test/test_algebraic.jl:38-69 (literal arrays in the test
files; nothing is executed).  Each entry: name -> {"shape": [...], "data": flat column-major
list (the Julia memory order), "source": "file:line", "call": the Julia call it pins}.
"""
import json
import os
import re

REF = "/root/reference/test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")

NAMES = {
    ("runtests.jl", 0): "fem1d_3nodes_p1",
    ("runtests.jl", 1): "fem2d_P2_quickstart_p1",
    ("runtests.jl", 2): "spectral1d_n5_p1",
    ("runtests.jl", 3): "spectral2d_n5_p1",
    ("runtests.jl", 4): "parabolic_fem1d_3nodes_h0.5_p1",
    ("runtests.jl", 5): "parabolic_fem2d_P2_h0.5_p1",
    ("runtests.jl", 6): "parabolic_spectral1d_n4_h0.5_p1",
    ("runtests.jl", 7): "parabolic_spectral2d_n4_h0.5_p1",
    ("test_algebraic.jl", 0): "fem1d_5nodes_p1",
    ("test_algebraic.jl", 1): "fem1d_5nodes_p1.5",
    ("test_algebraic.jl", 2): "fem2d_P1_L2_p1",
    ("test_algebraic.jl", 3): "fem2d_P1_L2_p1.5",
    ("test_algebraic.jl", 4): "fem2d_P2_L2_p1",
    ("test_algebraic.jl", 5): "fem2d_P2_L2_p1.5",
    ("test_algebraic.jl", 6): "fem3d_k1_L2_p1",
    ("test_algebraic.jl", 7): "fem3d_k1_L2_p1.5",
}


def parse_reshape(line):
    m = re.search(r"reshape\((?:Float64)?\[(.*?)\]\s*,\s*\(:\s*,\s*(\d+)\)\)", line)
    vals = [float(v) for v in m.group(1).split(",")]
    nc = int(m.group(2))
    return [len(vals) // nc, nc], vals


def parse_3d(line):
    body = line[line.index("[") + 1: line.rindex("]")]
    slabs = body.split(";;;")
    mats = []
    for s in slabs:
        rows = [[float(v) for v in r.split()] for r in s.split(";") if r.strip()]
        mats.append(rows)
    nr, nc, ns = len(mats[0]), len(mats[0][0]), len(mats)
    flat = [mats[k][i][j] for k in range(ns) for j in range(nc) for i in range(nr)]
    return [nr, nc, ns], flat


def main():
    out = {}
    for fname in ("runtests.jl", "test_algebraic.jl"):
        lines = open(os.path.join(REF, fname)).read().split("\n")
        k = 0
        for ln, line in enumerate(lines, 1):
            s = line.strip()
            if not s.startswith("z = "):
                continue
            if "reshape(" in s:
                shape, data = parse_reshape(s)
            elif ";;;" in s:
                shape, data = parse_3d(s)
            else:
                continue
            call = lines[ln].strip()
            name = NAMES[(fname, k)]
            out[name] = {"shape": shape, "data": data, "source": "test/%s:%d" % (fname, ln),
                         "call": call}
            k += 1
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0)
    for k, v in out.items():
        print(k, v["shape"], v["source"], v["call"][:70])


if __name__ == "__main__":
    main()
