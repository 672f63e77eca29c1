"""Deterministic synthetic checkpoints for parity tests.  TEST INFRASTRUCTURE (oracle).

The trained checkpoints of the reference zoo (perspectivefields.py:86-118) are not available offline,
so parity is checked on synthetic weights that BOTH sides load from the same ``{"model": state_dict}``
file through the normal hub-cache path (perspectivefields.py:178-192).

The values are drawn per key from ``numpy.random.RandomState(crc32(key) ^ seed)`` -- numpy's legacy
generator is bit-stable across platforms -- and scaled so that every layer matters: unit-gain linear
maps, non-zero biases, perturbed LayerNorm / BatchNorm affine parameters and statistics, O(0.1..0.5)
ConvNeXt layer-scale (the reference's 1e-6 init would hide the ParamNet blocks from any test), and a
ParamNet head biased towards plausible camera parameters so that general_vfov -> focal is solvable.
"""
import math
import zlib

import numpy as np
import torch

from .schema import state_dict_schema


def _rs(key, seed):
    return np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def synth_tensor(key, shape, seed=0, gravity_bias=(0.30, -0.80), gravity_gain=0.08):
    rs = _rs(key, seed)
    if key.endswith("num_batches_tracked"):
        return torch.tensor(1000, dtype=torch.int64)
    n = rs.standard_normal(shape).astype(np.float32) if len(shape) else None
    leaf = key.rsplit(".", 1)[1]
    mod = key.rsplit(".", 1)[0]
    modleaf = mod.rsplit(".", 1)[-1]

    if leaf == "running_var":
        t = rs.uniform(0.5, 1.5, shape).astype(np.float32)
    elif leaf == "running_mean":
        t = 0.2 * n
    elif leaf == "gamma":  # ConvNeXt layer scale
        t = rs.uniform(0.1, 0.5, shape).astype(np.float32)
    elif len(shape) == 1 and leaf == "weight":  # LayerNorm / BatchNorm scale
        t = 1.0 + 0.1 * n
    elif leaf == "bias":
        if mod == "param_net.backbone.head":
            t = np.array([0.05, -0.10, 0.60, 0.05, -0.05], dtype=np.float32)
        elif modleaf == "linear_pred_gravity" and shape[0] == 2:
            # a trained regression head emits near-unit up-vectors; keep |v| away from 0 so that the
            # F.normalize that follows is as well conditioned as it is in deployment
            t = np.array(gravity_bias, dtype=np.float32)
        else:
            t = 0.1 * n
    else:
        fan_in = int(np.prod(shape[1:]))
        gain = 1.0
        if key in ("ll_enc.conv1.weight", "backbone.patch_embed1.proj.weight"):
            gain = 1.0 / 74.0  # inputs are mean-subtracted bytes, std ~ 74
        elif "persformer_heads" in key and ("resConfUnit" in key or "conv_fuse" in key):
            gain = math.sqrt(2.0)  # convs that follow a ReLU
        elif modleaf == "linear_pred_latitude" and shape[0] == 1:
            gain = 0.04  # keep sin(latitude) mostly inside (-1, 1); calibrated in tests/golden/make_golden.py
        elif modleaf == "linear_pred_gravity" and shape[0] == 2:
            gain = gravity_gain
        elif mod == "param_net.backbone.head":
            gain = 0.3
        t = n * (gain / math.sqrt(fan_in))
    return torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))


def synth_state_dict(version, seed=0, gravity_bias=(0.30, -0.80), gravity_gain=0.08):
    """``OrderedDict``-like dict with the reference's key names and shapes.

    ``gravity_bias=(0, 0)`` with a larger ``gravity_gain`` gives a regression gravity head whose output is driven by the weights
    alone: the normalised field then turns through all directions and |v_raw| comes close to 0 at some pixels (the parity tests
    mask those), instead of the near-constant field of the default checkpoint."""
    return {k: synth_tensor(k, s, seed, gravity_bias, gravity_gain) for k, s in state_dict_schema(version)}


def synth_images(n, h, w, seed=0):
    """Synthetic BGR uint8 images: i.i.d. uniform bytes, drawn sequentially."""
    rs = np.random.RandomState(seed)
    return [rs.randint(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(n)]


def smooth_images(n, h, w, seed=0):
    """Smoother synthetic images (low-frequency sinusoid mixtures + noise): closer to photographs, so the
    antialiased resize and the first convolutions see non-trivial structure."""
    rs = np.random.RandomState(seed + 7919)
    out = []
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    for _ in range(n):
        img = np.zeros((h, w, 3), np.float64)
        for c in range(3):
            acc = np.zeros((h, w))
            for _k in range(4):
                fx, fy = rs.uniform(-0.03, 0.03, 2)
                acc += rs.uniform(20, 60) * np.sin(2 * np.pi * (fx * xx + fy * yy) + rs.uniform(0, 6.28))
            img[:, :, c] = 128 + acc + rs.normal(0, 6, (h, w))
        out.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    return out
