"""CPU, build container only: the oracle against the imported, unmodified reference (skipped where
/root/reference is absent, e.g. on the GPU box).  This is the live version of the golden fixtures."""
import os
import tempfile

import pytest
import torch

from oracle import model as om
from oracle import weights_gen as wg
from oracle.ref_shim import reference_available
from oracle.schema import state_dict_schema
from oracle.variants import VARIANTS

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def p2d():
    th = tempfile.mkdtemp(prefix="pf_ref_")
    os.environ["TORCH_HOME"] = th
    os.makedirs(os.path.join(th, "hub", "checkpoints"), exist_ok=True)
    from oracle.ref_shim import load_reference

    return load_reference(), th


@pytest.mark.parametrize("version", list(VARIANTS))
def test_schema_matches_reference(p2d, version):
    mod, th = p2d
    sd = {k: torch.zeros(s) for k, s in state_dict_schema(version)}
    torch.save({"model": sd}, os.path.join(th, "hub", "checkpoints", VARIANTS[version]["ckpt"]))
    ref_sd = mod.PerspectiveFields(version).state_dict()
    assert list(ref_sd.keys()) == [k for k, _ in state_dict_schema(version)]
    for k, s in state_dict_schema(version):
        assert tuple(ref_sd[k].shape) == tuple(s), k


def test_live_outputs_match(p2d):
    mod, th = p2d
    version = "PersNet_Paramnet-GSV-uncentered"
    sd = wg.synth_state_dict(version, 3)
    torch.save({"model": sd}, os.path.join(th, "hub", "checkpoints", VARIANTS[version]["ckpt"]))
    model = mod.PerspectiveFields(version).eval()
    imgs = wg.smooth_images(1, 300, 420, 5)
    ref = model.inference_batch(imgs)
    ora = om.inference_batch(sd, version, imgs)
    assert list(ref[0].keys()) == list(ora[0].keys())
    for k, v in ref[0].items():
        if isinstance(v, str):
            continue
        err = ((v - ora[0][k]).abs().max() / v.abs().max().clamp_min(1e-30)).item()
        assert err < 1e-4, (k, err)


def test_float_input_branch_matches_reference(p2d):
    """perspectivefields.py:47-66: non-uint8 images go through F.interpolate instead of PIL (pins oracle.model.inference_float /
    resize_float, which tests/test_gpu_forward.py uses as the referee for the CUDA float branch)."""
    import numpy as np

    mod, th = p2d
    version = "Paramnet-360Cities-edina-centered"
    sd = wg.synth_state_dict(version, 0)
    torch.save({"model": sd}, os.path.join(th, "hub", "checkpoints", VARIANTS[version]["ckpt"]))
    model = mod.PerspectiveFields(version).eval()
    img = wg.smooth_images(1, 200, 260, 9)[0].astype(np.float32) + 0.25
    assert np.array_equal(model.aug.apply_image(img), om.resize_float(img, 320, 320))
    ref = model.inference(img)
    ora = om.inference_float(sd, version, img)
    for k, v in ref.items():
        if isinstance(v, str):
            continue
        err = ((v - ora[k]).abs().max() / v.abs().max().clamp_min(1e-30)).item()
        assert err < 1e-4, (k, err)


def test_yaml_configuration_matches_reference(p2d):
    """perspectivefields_b200/config/*.yaml (defaults + per-variant overrides, parsed with PyYAML) give every inference-relevant
    field the value the reference's yacs tree has after merge_from_file (perspectivefields.py:124-131)."""
    from perspectivefields_b200 import variants as V

    mod, th = p2d
    for version in VARIANTS:
        sd = {k: torch.zeros(s) for k, s in state_dict_schema(version)}
        torch.save({"model": sd}, os.path.join(th, "hub", "checkpoints", VARIANTS[version]["ckpt"]))
        ref = mod.PerspectiveFields(version).cfg
        mine = V.make_cfg(version)

        def walk(a, b, path):
            for k, v in a.items():
                assert k in b, (version, path + k)
                if isinstance(v, dict):
                    walk(v, b[k], path + k + ".")
                else:
                    rv = b[k]
                    rv = list(rv) if isinstance(rv, (list, tuple)) else rv
                    assert rv == v, (version, path + k, rv, v)
        walk(mine, ref, "")
        assert V.model_zoo[version] == mod.perspectivefields.model_zoo[version]
