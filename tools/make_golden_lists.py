#!/usr/bin/env python3
"""Regenerates tests/golden/tiger256_{fills,tiles,zbuffer}.npy from the CPU oracle (oracle/pf_oracle.cpp).

The reference holds no golden vectors for tiles, fills, backdrops or alpha tiles and cannot
be built in this image (no Rust toolchain), so these files pin the ORACLE, not the reference: parity against
the real Rust tiler stays unpinned until tools/reference_dump/ has been run on a machine with cargo and its
output compared with tools/diff_lists.py.

Usage: python tools/make_golden_lists.py [--check]   (--check: compare with the committed files, write nothing)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pathfinder_b200 import scenes  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    check = "--check" in sys.argv[1:]
    flat, xf = scenes.tiger(256)
    built = H.oracle_build(flat, xf)
    out = {"fills": built.fills, "tiles": built.tiles, "zbuffer": built.z_buffer}
    ok = True
    for name, arr in out.items():
        path = os.path.join(ROOT, "tests", "golden", f"tiger256_{name}.npy")
        if check:
            same = os.path.exists(path) and np.array_equal(np.load(path), arr)
            print(f"{name}: {'identical' if same else 'DIFFERENT'} ({len(arr)} records)")
            ok &= same
        else:
            np.save(path, arr)
            print(f"wrote {path} ({len(arr)} records)")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
