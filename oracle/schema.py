"""Checkpoint schema ("model" state_dict keys, shapes) of the reference.  TEST INFRASTRUCTURE (oracle).

Restates the module tree built by perspective2d/perspectivefields.py:122-163:
``backbone`` = mit_b3 (mix_transformers.py:252-447,511-524), ``ll_enc`` (perspectivefields.py:70-83),
``persformer_heads.{gravity,latitude}_head`` (gravity_head.py:56-117, latitude_head.py:56-118),
``param_net.backbone`` = ConvNeXt-T with 5 outputs (convnext.py:62-130, param_network.py:41-43,178-181).
Checked against the imported reference in tests/test_oracle_vs_reference.py.
"""
from .variants import CNX_DEPTHS, CNX_DIMS, HEAD_EMBED, MIT_DEPTHS, MIT_DIMS, MIT_SR, VARIANTS


def _wb(out, prefix, wshape):
    out.append((prefix + ".weight", tuple(wshape)))
    out.append((prefix + ".bias", (wshape[0],)))


def _backbone(out):
    p = "backbone."
    cin = 3
    for s, c in enumerate(MIT_DIMS):
        k = 7 if s == 0 else 3
        _wb(out, f"{p}patch_embed{s + 1}.proj", (c, cin, k, k))
        _wb(out, f"{p}patch_embed{s + 1}.norm", (c,))
        cin = c
    for s, c in enumerate(MIT_DIMS):
        for i in range(MIT_DEPTHS[s]):
            b = f"{p}block{s + 1}.{i}."
            _wb(out, b + "norm1", (c,))
            _wb(out, b + "attn.q", (c, c))
            _wb(out, b + "attn.kv", (2 * c, c))
            _wb(out, b + "attn.proj", (c, c))
            if MIT_SR[s] > 1:
                _wb(out, b + "attn.sr", (c, c, MIT_SR[s], MIT_SR[s]))
                _wb(out, b + "attn.norm", (c,))
            _wb(out, b + "norm2", (c,))
            _wb(out, b + "mlp.fc1", (4 * c, c))
            _wb(out, b + "mlp.dwconv.dwconv", (4 * c, 1, 3, 3))
            _wb(out, b + "mlp.fc2", (c, 4 * c))
        _wb(out, f"{p}norm{s + 1}", (c,))


def _ll_enc(out):
    out.append(("ll_enc.conv1.weight", (64, 3, 7, 7)))
    for n in ("weight", "bias", "running_mean", "running_var"):
        out.append(("ll_enc.bn1." + n, (64,)))
    out.append(("ll_enc.bn1.num_batches_tracked", ()))


def _head(out, name, pred_name, ncls):
    p = f"persformer_heads.{name}."
    for lvl in (4, 3, 2, 1):
        _wb(out, f"{p}linear_c{lvl}.proj", (HEAD_EMBED, MIT_DIMS[lvl - 1]))
    for lvl in (4, 3, 2, 1):
        _wb(out, f"{p}linear_c{lvl}_proc", (256, HEAD_EMBED, 3, 3))
    for f in (1, 2, 3, 4):
        units = ("resConfUnit2",) if f == 4 else ("resConfUnit1", "resConfUnit2")
        for u in units:
            for cv in ("conv1", "conv2"):
                _wb(out, f"{p}fusion{f}.{u}.{cv}", (256, 256, 3, 3))
    _wb(out, p + "conv_fuse_conv0.conv", (64, 320, 3, 3))
    _wb(out, p + "conv_fuse_conv1.conv", (32, 64, 3, 3))
    _wb(out, p + pred_name, (ncls, 32, 1, 1))


def _param_net(out):
    p = "param_net.backbone."
    _wb(out, p + "downsample_layers.0.0", (CNX_DIMS[0], 3, 4, 4))
    _wb(out, p + "downsample_layers.0.1", (CNX_DIMS[0],))
    for k in (1, 2, 3):
        _wb(out, f"{p}downsample_layers.{k}.0", (CNX_DIMS[k - 1],))
        _wb(out, f"{p}downsample_layers.{k}.1", (CNX_DIMS[k], CNX_DIMS[k - 1], 2, 2))
    for s, c in enumerate(CNX_DIMS):
        for j in range(CNX_DEPTHS[s]):
            b = f"{p}stages.{s}.{j}."
            out.append((b + "gamma", (c,)))
            _wb(out, b + "dwconv", (c, 1, 7, 7))
            _wb(out, b + "norm", (c,))
            _wb(out, b + "pwconv1", (4 * c, c))
            _wb(out, b + "pwconv2", (c, 4 * c))
    _wb(out, p + "norm", (CNX_DIMS[3],))
    _wb(out, p + "head", (5, CNX_DIMS[3]))


def state_dict_schema(version):
    """Ordered ``[(key, shape)]`` of ``PerspectiveFields(version).state_dict()``."""
    v = VARIANTS[version]
    out = []
    _backbone(out)
    _ll_enc(out)
    _head(out, "gravity_head", "linear_pred_gravity", v["gravity_classes"])
    _head(out, "latitude_head", "linear_pred_latitude", v["latitude_classes"])
    if v["param_net"] is not None:
        _param_net(out)
    return out
