"""CPU: the integer restatement of Pillow's antialiased bilinear resize is bit-exact against Pillow itself
(the third-party library the reference calls at perspectivefields.py:42-44)."""
import numpy as np
import pytest
from PIL import Image

from oracle.pillow_resize import resize_bilinear_u8


@pytest.mark.parametrize("hw", [(480, 640), (240, 320), (768, 1024), (512, 512), (320, 320), (200, 300), (721, 900),
                                (33, 47), (320, 500), (500, 320), (1, 1), (1536, 2048)])
def test_bit_exact_vs_pillow(hw):
    h, w = hw
    rs = np.random.RandomState(h * 7 + w)
    img = rs.randint(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((320, 320), Image.BILINEAR))
    got = resize_bilinear_u8(img, 320, 320)
    assert got.dtype == np.uint8 and got.shape == (320, 320, 3)
    assert np.array_equal(got, ref)


def test_saturating_inputs():
    img = np.zeros((400, 600, 3), np.uint8)
    img[::2] = 255
    ref = np.asarray(Image.fromarray(img).resize((320, 320), Image.BILINEAR))
    assert np.array_equal(resize_bilinear_u8(img, 320, 320), ref)
