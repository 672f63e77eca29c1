"""CPU restatement of the visualisation hand-off of the reference's demo.  TEST INFRASTRUCTURE (oracle).

``resize_fix_aspect_ratio`` is a nested function of ``demo/demo.py:log_results`` (:30-51) and the arrow lattice lives inside
``draw_perspective_fields`` (perspective2d/utils/utils.py:190-200), which needs matplotlib: neither can be imported, so both
are restated here around the same third-party call the demo makes (``cv2.resize``, default INTER_LINEAR).  Parity of this file
is therefore pinned by inspection only ("parity unpinned" in the sense of the brief); the arithmetic is cv2's.
"""
import math

import cv2
import numpy as np


def resize_fix_aspect_ratio(field, target_width=None, target_height=None):
    """demo/demo.py:30-51 for the field dictionary: {"up": [2,H,W], "lati": [H,W]} float32 numpy -> resized copies."""
    height, width = field["lati"].shape
    if target_height is None:
        factor = target_width / width
    elif target_width is None:
        factor = target_height / height
    else:
        factor = max(target_width / width, target_height / height)
    if factor == target_width / width:
        target_height = int(height * factor)
    else:
        target_width = int(width * factor)
    out = {}
    for key in ("up", "lati"):
        tmp = field[key]
        transpose = tmp.ndim == 3
        if transpose:
            tmp = tmp.transpose(1, 2, 0)
        tmp = cv2.resize(np.ascontiguousarray(tmp), (target_width, target_height))
        if transpose:
            tmp = tmp.transpose(2, 0, 1)
        out[key] = tmp
    return out


def arrow_grid(up_hw2, density=10, arrow_inv_len=20):
    """perspective2d/utils/utils.py:190-200 (``up`` as [H, W, 2])."""
    im_h, im_w, _ = up_hw2.shape
    x, y = np.meshgrid(np.arange(0, im_w, im_w // density), np.arange(0, im_h, im_h // density))
    x, y = x.ravel(), y.ravel()
    arrow_len = np.sqrt(im_w ** 2 + im_h ** 2) // arrow_inv_len
    end = up_hw2[y, x, :] * arrow_len
    return x, y, end[:, 0], -end[:, 1]
