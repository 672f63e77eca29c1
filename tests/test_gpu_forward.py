"""GPU: the whole inference path through the product API (which calls libpf_b200.so through its C ABI) against the CPU
oracle on the same synthetic checkpoint and images, and against the golden fixtures made by the unmodified reference.

Tolerance (BASELINE.json north_star): 1e-3 relative fp32, metric = max|a - b| / max|b| per returned tensor.  The synthetic
gravity head emits vectors with |v| well away from 0 (oracle/weights_gen.py), so the normalised field is compared
directly.  argmax-decoded fields of the classification variant are compared only where the top-2 logit margin exceeds
the logit error bound (argmax is discontinuous)."""
import numpy as np
import pytest
import torch

import pf_test_util as U
from golden_util import compare_with_golden, golden_images
from oracle import model as om
from oracle import weights_gen as wg

pytestmark = pytest.mark.gpu
TOL = 1e-3

_models = {}


def model(version):
    if version not in _models:
        _models[version] = U.make_model(version)
    return _models[version]


def _check(out, ora, version, skip=()):
    worst = {}
    assert len(out) == len(ora)
    for o, r in zip(out, ora):
        assert list(o.keys()) == list(r.keys())
        for k, v in r.items():
            if isinstance(v, str):
                assert o[k] == v
                continue
            assert tuple(o[k].shape) == tuple(v.shape) and o[k].dtype == torch.float32, k
            if k in skip:
                continue
            if k == "pred_latitude_original" and v.abs().max() > 75:
                # degrees = asin(sin_lat): asin is not Lipschitz at +-1 (d/dx = 1/sqrt(1-x^2)), where the regression head
                # clamps; compare in the sine domain everywhere and in degrees away from the poles.
                a, b = o[k].detach().cpu().double(), v.double()
                e = U.rel_err(torch.sin(torch.deg2rad(a)), torch.sin(torch.deg2rad(b)))
                far = b.abs() < 75
                if far.any():
                    e = max(e, ((a - b).abs()[far].max() / b.abs().max()).item())
            else:
                e = U.rel_err(o[k], v)
            worst[k] = max(worst.get(k, 0.0), e)
            assert e < TOL, (version, k, e)
    return worst


@pytest.mark.parametrize("version", ["Paramnet-360Cities-edina-centered", "Paramnet-360Cities-edina-uncentered",
                                     "PersNet_Paramnet-GSV-uncentered", "PersNet_Paramnet-GSV-centered"])
def test_regression_variants_match_oracle_and_golden(version):
    m, sd = model(version)
    imgs = golden_images()
    out = m.inference_batch(imgs)
    assert all(o["pred_gravity"].is_cuda for o in out)
    print(version, _check(out, om.inference_batch(sd, version, imgs), version))
    print(version, "golden", compare_with_golden(version, out, tol=TOL, check_stats=True))


def test_classification_variant():
    version = "PersNet-360Cities"
    m, sd = model(version)
    imgs = golden_images()
    out = m.inference_batch(imgs)
    ora = om.inference_batch(sd, version, imgs)
    print(_check(out, ora, version, skip=("pred_gravity_original", "pred_latitude_original")))
    compare_with_golden(version, out, tol=TOL, skip_keys=("pred_gravity_original", "pred_latitude_original"))
    for o, r in zip(out, ora):
        # decoded fields: compare at 320x320 pixels whose argmax is unambiguous in the oracle logits
        for key, okey in (("pred_gravity", "pred_gravity_original"), ("pred_latitude", "pred_latitude_original")):
            top2 = r[key].topk(2, dim=0).values
            margin = top2[0] - top2[1]
            err = (o[key].cpu() - r[key]).abs().max()
            stable = margin > 4 * err
            assert stable.float().mean() > 0.9
            assert torch.equal(o[key].cpu().argmax(0)[stable], r[key].argmax(0)[stable])
        mism = (o["pred_latitude_original"].cpu() - r["pred_latitude_original"]).abs() > 1e-3 * 90
        assert mism.float().mean() < 0.02   # pixels next to a flipped (near-tie) argmax
        assert o["pred_latitude_original_mode"] == "deg"


def test_every_layer_tap_matches_oracle():
    version = "Paramnet-360Cities-edina-centered"
    m, sd = model(version)
    imgs = golden_images()
    m.debug_taps(True)
    try:
        m.inference_batch(imgs)
        taps = m.read_taps()
    finally:
        m.debug_taps(False)
    otaps = {}
    om.inference_batch(sd, version, imgs, otaps)
    checked = 0
    for name, t in taps.items():
        if name.startswith("mit."):
            ref = otaps[name]
        elif name == "ll" or name.startswith("cnx.s"):
            ref = otaps[name].permute(0, 2, 3, 1)
        elif name.startswith("head."):
            k = name[5:]
            ref = torch.cat([otaps["g." + k], otaps["l." + k]], 1).permute(0, 2, 3, 1)
        else:
            continue
        e = U.rel_err(t, ref.contiguous().reshape(-1))
        assert e < TOL, (name, e)
        checked += 1
    assert checked > 70


def test_mixed_sizes_identity_upscale_and_batch_invariance():
    version = "Paramnet-360Cities-edina-uncentered"
    m, sd = model(version)
    imgs = [wg.synth_images(1, 320, 320, 5)[0], wg.smooth_images(1, 33, 47, 6)[0], wg.smooth_images(1, 768, 1024, 7)[0],
            wg.synth_images(1, 240, 320, 8)[0]]
    keep = [im.copy() for im in imgs]
    out = m.inference_batch(imgs)
    assert all(np.array_equal(a, b) for a, b in zip(imgs, keep))  # inputs are not modified
    _check(out, om.inference_batch(sd, version, imgs), version)
    for i, im in enumerate(imgs):
        assert out[i]["pred_gravity_original"].shape == (2,) + im.shape[:2]
        assert out[i]["pred_latitude_original"].shape == im.shape[:2]
        single = m.inference(im)   # inference == inference_batch([x])[0]; sharding/batching never changes an image's result
        for k, v in single.items():
            if not isinstance(v, str):
                assert torch.equal(v, out[i][k]), (i, k)
    # a 320x320 input passes through the resize unchanged and the post-process resample is the identity
    assert torch.allclose(out[0]["pred_latitude_original"], torch.rad2deg(torch.asin(out[0]["pred_latitude"][0])), rtol=1e-6, atol=1e-5)
    assert torch.allclose(out[0]["pred_gravity_original"], out[0]["pred_gravity"], rtol=1e-6, atol=1e-6)


def test_forward_entry_with_preresized_float_images():
    version = "Paramnet-360Cities-edina-centered"
    m, sd = model(version)
    imgs = golden_images()
    inputs = [{"image": om.preprocess(im), "height": im.shape[0], "width": im.shape[1]} for im in imgs]
    a = m.forward(inputs)
    b = m.inference_batch(imgs)
    for x, y in zip(a, b):
        for k, v in x.items():
            if not isinstance(v, str):
                assert torch.equal(v, y[k]), k


def test_full_size_batch_properties():
    """BASELINE config C2 shape (batch 32, 640x480): size-independent properties + spot parity."""
    version = "Paramnet-360Cities-edina-centered"
    m, sd = model(version)
    imgs = wg.synth_images(32, 480, 640, 11)
    out = m.inference_batch(imgs)
    g = torch.stack([o["pred_gravity"] for o in out])
    go = torch.stack([o["pred_gravity_original"] for o in out])
    lat = torch.stack([o["pred_latitude"] for o in out])
    lo = torch.stack([o["pred_latitude_original"] for o in out])
    assert torch.isfinite(g).all() and torch.isfinite(go).all() and torch.isfinite(lo).all()
    assert (g.norm(dim=1) - 1).abs().max() < 1e-5 and (go.norm(dim=1) - 1).abs().max() < 1e-5
    assert lat.abs().max() <= 1 and lo.abs().max() <= 90.0001
    for i in (0, 17, 31):  # batch position never changes an image's result (what multi-GPU sharding relies on)
        single = m.inference(imgs[i])
        for k, v in single.items():
            if not isinstance(v, str):
                assert torch.equal(v, out[i][k]), (i, k)
    _check([out[5]], om.inference_batch(sd, version, [imgs[5]]), version)


def test_kernel_launches_are_counted():
    from perspectivefields_b200 import _native

    m, _ = model("Paramnet-360Cities-edina-uncentered")
    before = _native.lib().pf_kernel_launch_count()
    m.inference(wg.synth_images(1, 64, 64, 1)[0])
    assert _native.lib().pf_kernel_launch_count() - before > 300


@pytest.mark.parametrize("opts", [{"tma": 0}, {"tma": 0, "halo3x3": 0}, {"tma": 0, "tcgen05": 0}, {"attn_mma": 0, "stem_tc": 0}, {"phase_conv1": 0}])
def test_legacy_engines_end_to_end(opts):
    """The same forward on the earlier engines (fp32 activations split on the fly): register-staged tcgen05 kernels with
    / without the halo-tile 3x3 variant, and the warp-level HMMA kernel."""
    version = "Paramnet-360Cities-edina-centered"
    m, sd = model(version)
    imgs = golden_images()
    base = m.inference_batch(imgs)
    for k, v in opts.items():
        m.set_option(k, v)
    try:
        out = m.inference_batch(imgs)
    finally:
        for k in opts:
            m.set_option(k, 1)
    print(opts, _check(out, om.inference_batch(sd, version, imgs), version))
    compare_with_golden(version, out, tol=TOL)
    for a, b in zip(out, base):
        assert U.rel_err(a["pred_latitude"], b["pred_latitude"]) < 1e-4


def test_resolution_sweep_large_image_properties_and_parity():
    """BASELINE config C5 shape class: 2048x1536 and 1024x768 inputs (only the pre/post-processing bytes change)."""
    version = "PersNet_Paramnet-GSV-uncentered"
    m, sd = model(version)
    imgs = [wg.smooth_images(1, 1536, 2048, 21)[0], wg.synth_images(1, 768, 1024, 22)[0]]
    out = m.inference_batch(imgs)
    _check(out, om.inference_batch(sd, version, imgs), version)
    assert out[0]["pred_gravity_original"].shape == (2, 1536, 2048) and out[1]["pred_latitude_original"].shape == (768, 1024)
    assert (out[0]["pred_gravity_original"].norm(dim=0) - 1).abs().max() < 1e-5


def test_c3_shape_512x512_batch():
    """BASELINE config C3 shape: 512x512 inputs, uncentered ParamNet (64x64 nearest sub-sample), batch 8 here."""
    version = "Paramnet-360Cities-edina-uncentered"
    m, sd = model(version)
    imgs = wg.synth_images(8, 512, 512, 31)
    out = m.inference_batch(imgs)
    _check([out[3]], om.inference_batch(sd, version, [imgs[3]]), version)
    for o in out:
        assert torch.isfinite(o["pred_rel_focal"]) and o["pred_rel_focal"] > 0
