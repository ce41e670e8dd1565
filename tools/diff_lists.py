#!/usr/bin/env python3
"""Compares two list dumps (tools/dump_lists.py format) after the canonicalisation of SURVEY.md §8c:
alpha-tile ids are renumbered in order of first appearance in the fill list (SequentialExecutor order), fills
are compared as per-alpha-tile ordered lists, tiles as a set keyed by (path, y, x), z-buffers exactly.

  python tools/diff_lists.py ours.lists reference.lists
"""
import sys
from collections import OrderedDict


def read(path):
    fills, tiles, z = OrderedDict(), {}, None
    for line in open(path):
        w = line.split()
        if not w:
            continue
        if w[0] == "fill":
            fills.setdefault(int(w[5]), []).append(tuple(int(v) for v in w[1:5]))
        elif w[0] == "tile":
            tx, ty, alpha, pid, color, ctrl, backdrop = (int(v) for v in w[1:8])
            tiles[(pid, ty, tx)] = (alpha, color, ctrl, backdrop)
        elif w[0] == "z":
            z = tuple(int(v) for v in w[1:])
    return fills, tiles, z


def canonical(fills, tiles):
    rename = {old: new for new, old in enumerate(fills.keys())}
    cf = {rename[k]: v for k, v in fills.items()}
    ct = {}
    for key, (alpha, color, ctrl, backdrop) in tiles.items():
        solid = alpha == 0xFFFFFFFF or alpha not in rename
        ct[key] = (None if solid else rename[alpha], color, ctrl, backdrop)
    return cf, ct


def main():
    a, b = read(sys.argv[1]), read(sys.argv[2])
    fa, ta = canonical(a[0], a[1])
    fb, tb = canonical(b[0], b[1])
    problems = 0
    if len(fa) != len(fb):
        print(f"alpha tile count differs: {len(fa)} vs {len(fb)}")
        problems += 1
    for k in sorted(set(fa) | set(fb)):
        if fa.get(k) != fb.get(k):
            problems += 1
            if problems < 20:
                print(f"fills of alpha tile {k} differ: {fa.get(k)} vs {fb.get(k)}")
    for k in sorted(set(ta) | set(tb)):
        if ta.get(k) != tb.get(k):
            problems += 1
            if problems < 20:
                print(f"tile (path, y, x) = {k} differs: {ta.get(k)} vs {tb.get(k)}")
    if a[2] != b[2]:
        print("z-buffers differ")
        problems += 1
    print("IDENTICAL" if problems == 0 else f"{problems} differences")
    sys.exit(0 if problems == 0 else 1)


if __name__ == "__main__":
    main()
