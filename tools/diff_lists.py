#!/usr/bin/env python3
"""Compares two list dumps (tools/dump_lists.py format) after the canonicalisation of SURVEY.md §8c:
alpha-tile ids are renumbered in order of first appearance in the fill list (SequentialExecutor order), fills
are compared as per-alpha-tile ordered lists, tiles as a set keyed by (path, y, x), z-buffers exactly.

  python tools/diff_lists.py ours.lists reference.lists
"""
import sys
from collections import OrderedDict


def read(path):
    fills, tiles, z, clips = OrderedDict(), {}, None, []
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
        elif w[0] == "clip":
            clips.append(tuple(int(v) for v in w[1:5]))
    return fills, tiles, z, clips


def canonical(fills, tiles, clips=()):
    rename = {old: new for new, old in enumerate(fills.keys())}
    cf = {rename[k]: v for k, v in fills.items()}
    ct = {}
    for key, (alpha, color, ctrl, backdrop) in tiles.items():
        solid = alpha == 0xFFFFFFFF or alpha not in rename
        ct[key] = (None if solid else rename[alpha], color, ctrl, backdrop)
    cc = sorted((rename.get(d, d), db, rename.get(s, s), sb) for d, db, s, sb in clips)
    return cf, ct, cc


def main():
    a, b = read(sys.argv[1]), read(sys.argv[2])
    fa, ta, ca = canonical(a[0], a[1], a[3])
    fb, tb, cb = canonical(b[0], b[1], b[3])
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
    if ca != cb:
        print(f"clip records differ: {len(ca)} vs {len(cb)} records")
        problems += 1
    if a[2] != b[2]:
        print("z-buffers differ")
        problems += 1
    print("IDENTICAL" if problems == 0 else f"{problems} differences")
    sys.exit(0 if problems == 0 else 1)


if __name__ == "__main__":
    main()
