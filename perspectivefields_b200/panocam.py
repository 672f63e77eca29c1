"""Camera parameters -> dense perspective fields on the GPU (the step callers run right after the
inference path): drop-in for the two static methods of the reference's ``perspective2d.utils.panocam.PanoCam`` that turn
ParamNet's output into an up-vector field and a latitude map (utils/panocam.py:451-556; called from
utils/utils.py:367-385 and demo/demo.py:69-78).

Same names, argument order and meaning as the reference.  Differences: results are float32 CUDA tensors (the reference returns
float64 numpy arrays), and ``camera_fields`` evaluates a whole batch (images may differ in size) in one launch per 24 images.
There is no CPU path: the functions raise when the CUDA library or a CUDA device is missing.
"""
import ctypes
import math

import numpy as np
import torch

from . import _native


def general_vfov_to_focal(rel_cx, rel_cy, h, gvfov, degree):
    """utils/utils.py:47-91 (SciPy ``fsolve`` there): relative focal length from the general vertical field of view, the
    angle between the rays through the top-centre and bottom-centre pixels, for an off-centre principal point.  Closed form
    (DESIGN.md section 4): with c = cos(gvfov), A = f^2 + cx^2 + cy^2 + h^2/4:  4 (c^2 - 1) A^2 + 4 h^2 A - h^2 (h^2 + 4 c^2 cy^2) = 0,
    root with sign(2A - h^2) = sign(c).  Scalars or arrays; float64."""
    cx, cy, g = np.asarray(rel_cx, np.float64), np.asarray(rel_cy, np.float64), np.asarray(gvfov, np.float64)
    if degree:
        g = np.radians(g)
    c = np.cos(g)
    h = float(h)
    # p^2 = f^2 + cx^2 + (cy + h/2)^2 = A + h cy,  q^2 = A - h cy,  cos(gvfov) = (p^2 + q^2 - h^2) / (2 p q)
    # => (2A - h^2)^2 = 4 c^2 (A^2 - h^2 cy^2), a quadratic in A; squaring adds the root of the supplementary angle, which
    #    the sign condition removes
    a2 = 4.0 * (c * c - 1.0)
    a1 = 4.0 * h * h
    a0 = -(h ** 4 + 4.0 * c * c * h * h * cy * cy)
    disc = np.sqrt(np.maximum(a1 * a1 - 4.0 * a2 * a0, 0.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        r1, r2 = (-a1 + disc) / (2.0 * a2), (-a1 - disc) / (2.0 * a2)
        pick = np.where(np.sign(2.0 * r1 - h * h) == np.sign(c), r1, r2)
        lin = -a0 / a1                                   # c^2 == 1 never happens for a real field of view; guard anyway
        A = np.where(np.abs(a2) < 1e-300, lin, pick)
        f2 = A - cx * cx - cy * cy - h * h / 4.0
        return np.sqrt(f2)


def camera_fields(focal_rel, heights, widths, elevation, roll, cx_rel, cy_rel, device=None, up=True, lat=True):
    """Batched ``get_up_general`` / ``get_lat_general``: every argument is a sequence of length n (radians for the angles).
    Returns (list of [H_i, W_i, 2] float32 tensors or None, list of [H_i, W_i] float32 tensors in degrees or None)."""
    L = _native.lib()
    if not torch.cuda.is_available():
        raise RuntimeError("perspectivefields_b200.panocam needs a CUDA device (there is no CPU path)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("perspectivefields_b200.panocam needs a CUDA device (there is no CPU path)")
    n = len(heights)
    cams = (_native.pf_camera * n)()
    up_off = lat_off = 0
    for i in range(n):
        h, w = int(heights[i]), int(widths[i])
        cams[i] = _native.pf_camera(h, w, float(focal_rel[i]), float(elevation[i]), float(roll[i]), float(cx_rel[i]), float(cy_rel[i]),
                                    up_off, lat_off)
        up_off += 2 * h * w
        lat_off += h * w
    with torch.cuda.device(dev):
        up_blob = torch.empty(up_off, dtype=torch.float32, device=dev) if up else None
        lat_blob = torch.empty(lat_off, dtype=torch.float32, device=dev) if lat else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        _native.check(L.pf_camera_fields(dev.index if dev.index is not None else torch.cuda.current_device(), cams, n,
                                         up_blob.data_ptr() if up else None, lat_blob.data_ptr() if lat else None, stream))
    ups = [up_blob[c.up_offset:c.up_offset + 2 * c.height * c.width].view(c.height, c.width, 2) for c in cams] if up else None
    lats = [lat_blob[c.lat_offset:c.lat_offset + c.height * c.width].view(c.height, c.width) for c in cams] if lat else None
    return ups, lats


class PanoCam:
    """The two field-synthesis static methods of ``perspective2d.utils.panocam.PanoCam`` (same signatures)."""

    @staticmethod
    def get_up_general(focal_rel, im_w, im_h, elevation, roll, cx_rel, cy_rel, device=None):
        """utils/panocam.py:451-513 -> float32 CUDA tensor [im_h, im_w, 2]."""
        return camera_fields([focal_rel], [im_h], [im_w], [elevation], [roll], [cx_rel], [cy_rel], device, up=True, lat=False)[0][0]

    @staticmethod
    def get_lat_general(focal_rel, im_w, im_h, elevation, roll, cx_rel, cy_rel, device=None):
        """utils/panocam.py:515-556 -> float32 CUDA tensor [im_h, im_w], degrees."""
        return camera_fields([focal_rel], [im_h], [im_w], [elevation], [roll], [cx_rel], [cy_rel], device, up=False, lat=True)[1][0]


def fields_from_predictions(preds, sizes, mode="deg", device=None):
    """The parameter -> field step of ``draw_from_r_p_f_cx_cy`` (utils/utils.py:359-385) for a list of ``inference`` results:
    ``roll, pitch, general vfov, rel_cx, rel_cy`` (degrees when mode == "deg") -> focal by ``general_vfov_to_focal(cx, cy, 1,
    vfov, False)`` -> (up fields, latitude maps in degrees) at the given (H, W) sizes."""
    if mode not in ("deg", "rad"):
        raise ValueError("Bad argument")
    val = lambda d, k: float(d[k].item() if hasattr(d[k], "item") else d[k])
    roll = [val(p, "pred_roll") for p in preds]
    pitch = [val(p, "pred_pitch") for p in preds]
    vfov = [val(p, "pred_general_vfov") for p in preds]
    cx = [val(p, "pred_rel_cx") for p in preds]
    cy = [val(p, "pred_rel_cy") for p in preds]
    if mode == "deg":
        roll, pitch, vfov = [math.radians(v) for v in roll], [math.radians(v) for v in pitch], [math.radians(v) for v in vfov]
    focal = general_vfov_to_focal(cx, cy, 1, vfov, False)
    return camera_fields(list(focal), [s[0] for s in sizes], [s[1] for s in sizes], pitch, roll, cx, cy, device)
