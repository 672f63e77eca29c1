"""GPU: pf_camera_fields (camera parameters -> up-vector field and latitude map) through the C ABI against
the oracle restatement of PanoCam.get_up_general / get_lat_general, on the reference-generated golden cases, on random
parameters, and batched with mixed image sizes.  float64 math, float32 results: tolerance 2e-6 on the unit vectors and
2e-5 degrees on the latitude (float32 rounding of values up to 90)."""
import os

import numpy as np
import pytest
import torch

from oracle import panocam as op
from perspectivefields_b200 import panocam as pc

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "panocam.npz"))
UP_TOL, LAT_TOL = 2e-6, 2e-5


def test_single_image_methods_match_reference_golden():
    for i, (f, w, h, el, roll, cx, cy) in enumerate(GOLD["cases"]):
        up = pc.PanoCam.get_up_general(f, int(w), int(h), el, roll, cx, cy)
        lat = pc.PanoCam.get_lat_general(f, int(w), int(h), el, roll, cx, cy)
        assert up.is_cuda and up.dtype == torch.float32 and tuple(up.shape) == (int(h), int(w), 2) and tuple(lat.shape) == (int(h), int(w))
        assert np.abs(up.cpu().numpy() - GOLD[f"up{i}"]).max() < UP_TOL
        assert np.abs(lat.cpu().numpy() - GOLD[f"lat{i}"]).max() < LAT_TOL


def test_batch_with_mixed_sizes_matches_oracle():
    rs = np.random.RandomState(11)
    n = 53                                                  # more than one launch chunk (24 images per launch)
    f = rs.uniform(0.3, 2.5, n); el = rs.uniform(-1.5, 1.5, n); roll = rs.uniform(-3.1, 3.1, n)
    cx, cy = rs.uniform(-0.3, 0.3, n), rs.uniform(-0.3, 0.3, n)
    hs, ws = rs.randint(1, 90, n), rs.randint(1, 120, n)
    el[5] = 0.0; el[17] = -0.0                              # the constant-field branch
    hs[3], ws[3] = 480, 640
    ups, lats = pc.camera_fields(f, hs, ws, el, roll, cx, cy)
    torch.cuda.synchronize()
    for i in range(n):
        ru = op.get_up_general(f[i], int(ws[i]), int(hs[i]), el[i], roll[i], cx[i], cy[i])
        rl = op.get_lat_general(f[i], int(ws[i]), int(hs[i]), el[i], roll[i], cx[i], cy[i])
        assert np.abs(ups[i].cpu().numpy() - ru).max() < UP_TOL, i
        assert np.abs(lats[i].cpu().numpy() - rl).max() < LAT_TOL, i


def test_fields_from_predictions_full_size_properties():
    """640x480 fields from ParamNet-style predictions (degrees): unit up-vectors; latitude within [-90, 90]; a level camera
    (pitch = roll = 0, centred) has latitude 0 on the middle row pair and an up field of exactly (0, -1)."""
    preds = [{"pred_roll": torch.tensor(3.0), "pred_pitch": torch.tensor(-12.0), "pred_general_vfov": torch.tensor(55.0),
              "pred_rel_cx": torch.tensor(0.05), "pred_rel_cy": torch.tensor(-0.02)},
             {"pred_roll": 0.0, "pred_pitch": 0.0, "pred_general_vfov": 60.0, "pred_rel_cx": 0.0, "pred_rel_cy": 0.0}]
    ups, lats = pc.fields_from_predictions(preds, [(480, 640), (480, 640)], "deg")
    for u, l in zip(ups, lats):
        assert torch.allclose(u.norm(dim=2), torch.ones_like(l), atol=1e-6)
        assert l.abs().max() <= 90.0
    assert torch.equal(ups[1], torch.tensor([0.0, -1.0], device=ups[1].device).expand(480, 640, 2))
    assert lats[1][239:241].abs().max() < 0.15 and torch.allclose(lats[1][239], -lats[1][240], atol=1e-5)
    f = pc.general_vfov_to_focal([0.05], [-0.02], 1, np.radians([55.0]), False)
    ref = op.get_lat_general(float(f[0]), 640, 480, np.radians(-12.0), np.radians(3.0), 0.05, -0.02)
    assert np.abs(lats[0].cpu().numpy() - ref).max() < LAT_TOL
