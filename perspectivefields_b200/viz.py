"""Visualisation hand-off on the GPU: what the reference's demo does to the predicted fields between the path
and matplotlib -- ``resize_fix_aspect_ratio`` (demo/demo.py:30-51: ``cv2.resize`` of the up field and the latitude map to a
640-pixel-wide canvas) and the arrow grid of ``draw_perspective_fields`` / ``draw_up_field`` (utils/utils.py:190-200,
:236-247: ``up[y, x] * arrow_len`` on a ``density`` x ``density`` lattice) -- evaluated on the device, so that only the
canvas-sized latitude map (for the contour plot) and a few hundred arrow coordinates cross PCIe instead of the full-resolution
fields (37 MB per 2048x1536 image).

The resampling is ``pf_op_resize_f32`` (bilinear, pixel-centre aligned, no antialias: the sampling positions of
``cv2.resize(..., INTER_LINEAR)`` and of ``F.interpolate(align_corners=False)`` coincide).  No CPU path.
"""
import math

import numpy as np
import torch

from . import _native


def _resize_plane(t, th, tw):
    """[H, W] float32 CUDA tensor -> [th, tw]."""
    L = _native.lib()
    t = t.contiguous()
    out = torch.empty((th, tw), dtype=torch.float32, device=t.device)
    stream = torch.cuda.current_stream(t.device).cuda_stream
    _native.check(L.pf_op_resize_f32(t.data_ptr(), t.shape[0], t.shape[1], 1, th, tw, out.data_ptr(), stream))
    return out


def target_size(height, width, target_width=None, target_height=None):
    """demo/demo.py:30-40: the canvas size ``resize_fix_aspect_ratio`` picks."""
    if target_width is None and target_height is None:
        raise ValueError("target_width or target_height must be given")
    if target_height is None:
        factor = target_width / width
    elif target_width is None:
        factor = target_height / height
    else:
        factor = max(target_width / width, target_height / height)
    if target_width is not None and factor == target_width / width:
        target_height = int(height * factor)
    else:
        target_width = int(width * factor)
    return target_height, target_width


def resize_fields(up, lati, target_width=640, target_height=None):
    """``resize_fix_aspect_ratio`` for the fields (demo/demo.py:41-51): ``up`` [2, H, W] and ``lati`` [H, W] CUDA tensors ->
    ([2, th, tw], [th, tw]) CUDA tensors."""
    if up.device.type != "cuda":
        raise RuntimeError("perspectivefields_b200.viz needs CUDA tensors (there is no CPU path)")
    h, w = lati.shape
    th, tw = target_size(h, w, target_width, target_height)
    with torch.cuda.device(up.device):
        up_r = torch.stack([_resize_plane(up[0].float(), th, tw), _resize_plane(up[1].float(), th, tw)])
        lat_r = _resize_plane(lati.float(), th, tw)
    return up_r, lat_r


def arrow_grid(up, density=10, arrow_inv_len=20):
    """utils/utils.py:190-200: lattice ``x = arange(0, w, w // density)``, ``y = arange(0, h, h // density)``; arrows
    ``up[:, y, x] * (sqrt(w^2 + h^2) // arrow_inv_len)``, drawn with (u, -v).  ``up``: [2, H, W] (CUDA).  Returns host arrays
    x, y (int64) and u, v (float32, v already negated as ``draw_arrow`` receives it)."""
    _, h, w = up.shape
    xs = torch.arange(0, w, w // density, device=up.device)
    ys = torch.arange(0, h, h // density, device=up.device)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")          # np.meshgrid(x, y) ravel order: y outer, x inner
    x, y = xx.reshape(-1), yy.reshape(-1)
    arrow_len = math.sqrt(w ** 2 + h ** 2) // arrow_inv_len
    end = up[:, y, x] * arrow_len
    return x.cpu().numpy(), y.cpu().numpy(), end[0].cpu().numpy(), (-end[1]).cpu().numpy()


def handoff(pred, target_width=640, density=10, arrow_inv_len=20):
    """Everything ``demo.log_results`` needs from one prediction to draw ``perspective_pred`` (demo/demo.py:53-62): the latitude
    map in RADIANS on the canvas (host float32, for ``draw_lati``'s contours) and the arrow lattice of the up field."""
    up_r, lat_r = resize_fields(pred["pred_gravity_original"], pred["pred_latitude_original"], target_width)
    x, y, u, v = arrow_grid(up_r, density, arrow_inv_len)
    return {"latitude_rad": torch.deg2rad(lat_r).cpu().numpy(), "arrow_x": x, "arrow_y": y, "arrow_u": u, "arrow_v": v,
            "canvas_hw": tuple(lat_r.shape)}
