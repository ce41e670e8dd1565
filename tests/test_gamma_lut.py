"""The gamma-correction table of the text filter: pathfinder_b200/gamma_lut.py restates
utils/gamma-lut and is pinned to the reference's shipped resources/textures/gamma-lut.png — one of the few golden
vectors the reference holds near this path — and to the reference's own unit test of the table builder."""
import hashlib
import os

import numpy as np
import pytest

from pathfinder_b200 import gamma_lut

REF_PNG = "/root/reference/resources/textures/gamma-lut.png"


def test_generated_lut_checksum():
    lut = gamma_lut.generate()
    assert lut.shape == (8, 256) and lut.dtype == np.uint8
    assert hashlib.sha256(lut.tobytes()).hexdigest() == gamma_lut.SHA256


@pytest.mark.skipif(not os.path.exists(REF_PNG), reason="reference checkout not present")
def test_generated_lut_matches_reference_png():
    from PIL import Image
    ref = np.array(Image.open(REF_PNG))
    assert ref.shape == (8, 256)
    assert np.array_equal(gamma_lut.generate(), ref)


def test_scale255_replicates_bits():
    assert [gamma_lut.scale255(3, i) for i in range(8)] == [0x00, 0x24, 0x49, 0x6D, 0x92, 0xB6, 0xDB, 0xFF]
    assert gamma_lut.scale255(8, 0xAB) == 0xAB and gamma_lut.scale255(1, 1) == 0xFF


def test_lut_structure():
    lut = gamma_lut.generate().astype(np.int32)
    assert (lut[:, 0] == 0).all() and (lut[:, 255] == 255).all()   # no coverage / full coverage are fixed points
    assert (np.diff(lut, axis=1) >= 0).all()                        # monotonic in coverage
    # dark text on a light ground is thinned, light text on a dark ground is thickened
    assert (lut[0, 1:255] <= np.arange(1, 255)).all() and (lut[7, 1:255] >= np.arange(1, 255)).all()
    assert (np.diff(lut[:, 128]) > 0).all()


def test_reference_unit_test_gamma():
    """gamma_lut.rs:316-353 restated: with paint and device space Gamma(2.0), blending src over dst with the
    preblended alpha differs from the linear-light blend by at most 33/255, for src 131..255, every dst and alpha."""
    to_luma, from_luma = gamma_lut.gamma_to_luma(2.0)
    g = np.float32(2.0)
    dst = np.arange(256, dtype=np.uint32)[:, None]
    alpha = np.arange(256, dtype=np.uint32)[None, :]
    worst = 0
    for src in range(131, 256):
        table = gamma_lut.build_gamma_correcting_lut(src, 0.0, to_luma, from_luma).astype(np.uint32)
        preblend = table[None, :]
        preblend_result = (src * preblend + dst * (255 - preblend)) // 255
        f = np.float32
        lin_dst = np.power(dst.astype(f) / f(255.0), g) * f(255.0)
        lin_src = np.power(f(src) / f(255.0), g) * f(255.0)
        over = (lin_src * alpha.astype(f) + lin_dst * (f(255.0) - alpha.astype(f))) / f(255.0)
        true_result = (np.power(over / f(255.0), f(1.0) / g) * f(255.0)).astype(np.uint32)
        diff = np.abs(preblend_result.astype(np.int64) - true_result.astype(np.int64)).max()
        worst = max(worst, int(diff))
    assert worst <= 33, worst
