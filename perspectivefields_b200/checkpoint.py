"""Checkpoint drop-in contract: key names / shapes of the reference's ``{"model": state_dict}`` files and the loader.

Follows perspective2d/perspectivefields.py:178-192 (``torch.hub.load_state_dict_from_url(url, map_location=cpu)`` into
``$TORCH_HOME/hub/checkpoints/<basename>``, then ``load_state_dict(ckpt["model"], strict=False)``) and the module tree
of perspectivefields.py:122-163 (SURVEY.md appendix A lists the resulting keys).
"""
import math

import torch

from .variants import CNX_DEPTHS, CNX_DIMS, HEAD_EMBED, MIT_DEPTHS, MIT_DIMS, MIT_SR, VARIANTS


def checkpoint_schema(version):
    """Ordered list of (key, shape) of ``perspective2d.PerspectiveFields(version).state_dict()``."""
    v = VARIANTS[version]
    keys = []

    def wb(prefix, *wshape):
        keys.append((prefix + ".weight", tuple(wshape)))
        keys.append((prefix + ".bias", (wshape[0],)))

    cin = 3
    for s, c in enumerate(MIT_DIMS):
        k = 7 if s == 0 else 3
        wb(f"backbone.patch_embed{s + 1}.proj", c, cin, k, k)
        wb(f"backbone.patch_embed{s + 1}.norm", c)
        cin = c
    for s, c in enumerate(MIT_DIMS):
        for i in range(MIT_DEPTHS[s]):
            b = f"backbone.block{s + 1}.{i}."
            wb(b + "norm1", c)
            wb(b + "attn.q", c, c)
            wb(b + "attn.kv", 2 * c, c)
            wb(b + "attn.proj", c, c)
            if MIT_SR[s] > 1:
                wb(b + "attn.sr", c, c, MIT_SR[s], MIT_SR[s])
                wb(b + "attn.norm", c)
            wb(b + "norm2", c)
            wb(b + "mlp.fc1", 4 * c, c)
            wb(b + "mlp.dwconv.dwconv", 4 * c, 1, 3, 3)
            wb(b + "mlp.fc2", c, 4 * c)
        wb(f"backbone.norm{s + 1}", c)
    keys.append(("ll_enc.conv1.weight", (64, 3, 7, 7)))
    for n in ("weight", "bias", "running_mean", "running_var"):
        keys.append(("ll_enc.bn1." + n, (64,)))
    keys.append(("ll_enc.bn1.num_batches_tracked", ()))
    for head, pred, ncls in (("gravity_head", "linear_pred_gravity", v["gravity_classes"]),
                             ("latitude_head", "linear_pred_latitude", v["latitude_classes"])):
        p = f"persformer_heads.{head}."
        for lvl in (4, 3, 2, 1):
            wb(f"{p}linear_c{lvl}.proj", HEAD_EMBED, MIT_DIMS[lvl - 1])
        for lvl in (4, 3, 2, 1):
            wb(f"{p}linear_c{lvl}_proc", 256, HEAD_EMBED, 3, 3)
        for f in (1, 2, 3, 4):
            for u in ((2,) if f == 4 else (1, 2)):
                for c in (1, 2):
                    wb(f"{p}fusion{f}.resConfUnit{u}.conv{c}", 256, 256, 3, 3)
        wb(p + "conv_fuse_conv0.conv", 64, 320, 3, 3)
        wb(p + "conv_fuse_conv1.conv", 32, 64, 3, 3)
        wb(p + pred, ncls, 32, 1, 1)
    if v["param_net"] is not None:
        p = "param_net.backbone."
        wb(p + "downsample_layers.0.0", CNX_DIMS[0], 3, 4, 4)
        wb(p + "downsample_layers.0.1", CNX_DIMS[0])
        for k in (1, 2, 3):
            wb(f"{p}downsample_layers.{k}.0", CNX_DIMS[k - 1])
            wb(f"{p}downsample_layers.{k}.1", CNX_DIMS[k], CNX_DIMS[k - 1], 2, 2)
        for s, c in enumerate(CNX_DIMS):
            for j in range(CNX_DEPTHS[s]):
                b = f"{p}stages.{s}.{j}."
                keys.append((b + "gamma", (c,)))
                wb(b + "dwconv", c, 1, 7, 7)
                wb(b + "norm", c)
                wb(b + "pwconv1", 4 * c, c)
                wb(b + "pwconv2", c, 4 * c)
        wb(p + "norm", CNX_DIMS[3])
        wb(p + "head", 5, CNX_DIMS[3])
    return keys


def default_state(version, seed=0):
    """Deterministic stand-in for the reference's random module init (values a checkpoint does not cover keep these,
    as with ``strict=False`` in the reference): N(0, 0.02) weights, unit norms, zero biases, identity BN statistics."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in checkpoint_schema(version):
        leaf = k.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            sd[k] = torch.zeros((), dtype=torch.int64)
        elif leaf in ("bias", "running_mean"):
            sd[k] = torch.zeros(shape)
        elif leaf == "running_var" or (leaf == "weight" and len(shape) == 1):
            sd[k] = torch.ones(shape)
        elif leaf == "gamma":
            sd[k] = torch.full(shape, 1e-6)
        else:
            fan_in = int(math.prod(shape[1:]))
            sd[k] = torch.randn(shape, generator=g) * min(0.02, 1.0 / math.sqrt(fan_in))
    return sd


def load_zoo_checkpoint(url):
    """perspectivefields.py:181-184: hub cache first, download otherwise; always mapped to CPU."""
    return torch.hub.load_state_dict_from_url(url, map_location=torch.device("cpu"))
