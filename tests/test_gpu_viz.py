"""GPU: visualisation hand-off and the JPEG decode front-end (8f-2) against their host counterparts (cv2)."""
import numpy as np
import pytest
import torch

import pf_test_util as U
from oracle import viz as oviz
from oracle import weights_gen as wg

pytestmark = pytest.mark.gpu


def test_resize_fields_and_arrow_grid_match_the_demo_arithmetic():
    from perspectivefields_b200 import viz

    version = "Paramnet-360Cities-edina-centered"
    m, _ = U.make_model(version)
    img = wg.smooth_images(1, 768, 1024, 3)[0]
    pred = m.inference(img)
    h = viz.handoff(pred, target_width=640)
    field = {"up": pred["pred_gravity_original"].cpu().numpy(), "lati": pred["pred_latitude_original"].cpu().numpy()}
    ref = oviz.resize_fix_aspect_ratio(field, 640)
    assert h["canvas_hw"] == ref["lati"].shape == (480, 640)
    assert np.abs(np.degrees(h["latitude_rad"]) - ref["lati"]).max() < 1e-3          # degrees
    up_r, _ = viz.resize_fields(pred["pred_gravity_original"], pred["pred_latitude_original"], 640)
    assert np.abs(up_r.cpu().numpy() - ref["up"]).max() < 1e-5
    x, y, u, v = oviz.arrow_grid(ref["up"].transpose(1, 2, 0))
    assert np.array_equal(h["arrow_x"], x) and np.array_equal(h["arrow_y"], y)
    assert np.abs(h["arrow_u"] - u).max() < 1e-3 and np.abs(h["arrow_v"] - v).max() < 1e-3
    # a down-scaled and an up-scaled canvas with non-dyadic ratios
    for tw in (333, 1500):
        up_r, lat_r = viz.resize_fields(pred["pred_gravity_original"], pred["pred_latitude_original"], tw)
        ref = oviz.resize_fix_aspect_ratio(field, tw)
        assert tuple(lat_r.shape) == ref["lati"].shape
        assert np.abs(lat_r.cpu().numpy() - ref["lati"]).max() < 2e-3 and np.abs(up_r.cpu().numpy() - ref["up"]).max() < 2e-5


def test_jpeg_front_end_matches_cv2_decode():
    """decode_batch(JPEG bytes) vs cv2.imdecode (= what cv2.imread reads, demo/demo.py:151): same sizes, BGR order, pixels equal
    up to the decoders' rounding -- nvJPEG and libjpeg use different (both standard-conforming) IDCT / chroma up-sampling
    arithmetic, so a few grey levels of difference are expected (the numbers are printed).  Then the whole path on the decoded blob."""
    import cv2

    version = "Paramnet-360Cities-edina-uncentered"
    m, _ = U.make_model(version)
    imgs = wg.smooth_images(3, 360, 500, 11) + wg.smooth_images(1, 240, 320, 12)
    for quality, sampling in ((95, None), (100, getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_444", None))):
        params = [cv2.IMWRITE_JPEG_QUALITY, quality]
        if sampling is not None:
            params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sampling]
        jpegs = [cv2.imencode(".jpg", im, params)[1].tobytes() for im in imgs]
        dec = [cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR) for b in jpegs]
        blob, offsets, hs, ws = m.decode_batch(jpegs)
        torch.cuda.synchronize()
        for i, d in enumerate(dec):
            assert (hs[i], ws[i]) == d.shape[:2]
            g = blob[offsets[i]:offsets[i] + d.size].view(d.shape).cpu().numpy().astype(np.int32)
            diff = np.abs(g - d.astype(np.int32))
            swapped = np.abs(g[..., ::-1] - d.astype(np.int32))
            print(f"q{quality} image {i}: mean |diff| {diff.mean():.3f}, 99.9 % {np.percentile(diff, 99.9):.0f}, max {diff.max()}")
            # 4:2:0 streams: the two decoders up-sample chroma differently (libjpeg's "fancy" triangle filter vs nvJPEG), a few
            # grey levels on average; 4:4:4 streams differ by IDCT / colour-conversion rounding only
            lim_mean, lim_p = (3.5, 24) if sampling is None else (1.2, 5)
            assert diff.mean() < lim_mean and np.percentile(diff, 99.9) <= lim_p, (quality, i, diff.mean(), diff.max())
            assert diff.mean() < swapped.mean()          # channel order is BGR, as cv2 delivers
    a = m.inference_batch_encoded(jpegs)
    b = m.inference_batch(dec)
    assert len(a) == len(b) == 4
    for x, y in zip(a, b):
        assert list(x.keys()) == list(y.keys())
        for k, v in y.items():
            if not isinstance(v, str):
                # (the synthetic random-weight network amplifies the decoders' +-1..2 grey levels to a few percent: this is a
                #  plumbing check -- sizes, offsets, channel order --, not a kernel tolerance)
                assert tuple(x[k].shape) == tuple(v.shape) and U.rel_err(x[k], v) < 0.2, k
