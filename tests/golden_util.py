"""Helpers shared by the golden-fixture tests (oracle on CPU, CUDA path on the GPU box)."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def manifest():
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as f:
        return json.load(f)


def load_golden(version):
    m = manifest()
    return m, dict(np.load(os.path.join(GOLDEN_DIR, m["versions"][version]["file"])))


def golden_images():
    from oracle import weights_gen as wg

    m = manifest()
    return wg.synth_images(1, 480, 640, m["seed"]) + wg.smooth_images(1, 360, 500, m["seed"])


def rel_err(a, b):
    """max|a-b| / max|b|  -- the 'relative fp32 tolerance' metric of BASELINE.md (b = reference)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def compare_with_golden(version, results, tol, check_stats=True, skip_keys=()):
    """Compare a list of result dicts (one per golden image) with the stored reference outputs.
    Returns {key: worst relative error}."""
    m, g = load_golden(version)
    worst = {}
    for i, res in enumerate(results):
        assert list(res.keys()) == m["versions"][version]["keys"][i], (list(res.keys()), m["versions"][version]["keys"][i])
        for k, v in res.items():
            if isinstance(v, str):
                assert v == "deg"
                continue
            if k in skip_keys:
                continue
            v = v.detach().cpu().float()
            assert tuple(v.shape) == tuple(g[f"{i}/{k}/shape"]), (k, v.shape)
            st = m["logit_stride"] if (v.ndim == 3 and v.shape[0] > 3) else m["stride"]
            sub = v[..., ::st, ::st] if v.ndim >= 2 else v
            e = rel_err(sub, g[f"{i}/{k}"])
            worst[k] = max(worst.get(k, 0.0), e)
            assert e <= tol, f"{version} img{i} {k}: rel err {e:.3g} > {tol}"
            if check_stats:
                s, sa, n = g[f"{i}/{k}/stats"]
                assert abs(v.double().abs().sum().item() - sa) <= tol * max(sa, 1e-30) * 4, (k, "abs-sum checksum")
    return worst
