import hashlib
import os

import numpy as np

from pathfinder_b200 import area_lut

REF_PNG = "/root/reference/resources/textures/area-lut.png"


def test_generated_lut_checksum():
    lut = area_lut.generate()
    assert lut.shape == (256, 256, 4) and lut.dtype == np.uint8
    assert hashlib.sha256(lut.tobytes()).hexdigest() == area_lut.SHA256


def test_generated_lut_matches_reference_png():
    """Pins the restated generator (utils/area-lut/src/main.rs) to the reference's shipped texture
    whenever the reference checkout is present (it is not on the GPU box)."""
    if not os.path.exists(REF_PNG):
        import pytest
        pytest.skip("reference checkout not present")
    from PIL import Image
    ref = np.array(Image.open(REF_PNG))
    assert np.array_equal(area_lut.generate(), ref)


def test_lut_structure():
    lut = area_lut.generate()
    assert (lut[:, 0] == 255).all() and (lut[:, 255] == 0).all()  # main.rs:78-83
    # channel k is channel 0 shifted by 16 texels (one pixel row)
    for k in range(1, 4):
        assert np.array_equal(lut[:, 16 * k + 1:255, k], lut[:, 1:255 - 16 * k, 0])
    # horizontal line through the pixel centre covers half of it
    assert lut[0, 128, 0] == 128
