"""PyTorch-CPU fp32 restatement of the reference inference path.  TEST INFRASTRUCTURE (oracle).

Functional (state_dict in, tensors out); every function cites the reference code it follows.  Third-party
arithmetic: ATen CPU kernels through ``torch.nn.functional`` (the same calls the
reference makes), Pillow's resampler (restated in oracle/pillow_resize.py), SciPy ``fsolve``.
Pinned against the imported reference in tests/test_oracle_vs_reference.py and by tests/golden/.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .pillow_resize import resize_bilinear_u8
from .variants import (CNX_DEPTHS, CNX_DIMS, MIT_DEPTHS, MIT_DIMS, MIT_HEADS, MIT_SR, NET_H, NET_W, PIXEL_MEAN,
                       PIXEL_STD, VARIANTS)


# ----------------------------------------------------------------------------------------------- MiT-B3
def _attention(sd, p, x, H, W, heads, sr):
    """mix_transformers.py:108-141 (Attention.forward)."""
    B, N, C = x.shape
    d = C // heads
    q = F.linear(x, sd[p + "q.weight"], sd[p + "q.bias"]).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    if sr > 1:
        x_ = x.permute(0, 2, 1).reshape(B, C, H, W)
        x_ = F.conv2d(x_, sd[p + "sr.weight"], sd[p + "sr.bias"], stride=sr).reshape(B, C, -1).permute(0, 2, 1)
        x_ = F.layer_norm(x_, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)  # nn.LayerNorm default eps, :89
    else:
        x_ = x
    kv = F.linear(x_, sd[p + "kv.weight"], sd[p + "kv.bias"]).reshape(B, -1, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = (q @ k.transpose(-2, -1)) * (d ** -0.5)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def _mix_ffn(sd, p, x, H, W):
    """mix_transformers.py:49-56 (Mlp.forward) + :502-508 (DWConv.forward)."""
    B, N, _ = x.shape
    x = F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"])
    C4 = x.shape[-1]
    x = x.transpose(1, 2).contiguous().view(B, C4, H, W)
    x = F.conv2d(x, sd[p + "dwconv.dwconv.weight"], sd[p + "dwconv.dwconv.bias"], padding=1, groups=C4)
    x = x.flatten(2).transpose(1, 2)
    x = F.gelu(x)  # nn.GELU default = exact erf, :20
    return F.linear(x, sd[p + "fc2.weight"], sd[p + "fc2.bias"])


def mit_b3(sd, images, taps=None):
    """mix_transformers.py:449-485 (forward_features); block = :198-202; patch embed = :243-249."""
    p = "backbone."
    x = images
    outs = []
    for s in range(4):
        C = MIT_DIMS[s]
        k, st = (7, 4) if s == 0 else (3, 2)
        x = F.conv2d(x, sd[f"{p}patch_embed{s + 1}.proj.weight"], sd[f"{p}patch_embed{s + 1}.proj.bias"],
                     stride=st, padding=k // 2)
        B, _, H, W = x.shape
        x = x.flatten(2).transpose(1, 2)
        x = F.layer_norm(x, (C,), sd[f"{p}patch_embed{s + 1}.norm.weight"], sd[f"{p}patch_embed{s + 1}.norm.bias"], 1e-5)
        if taps is not None:
            taps[f"mit.s{s + 1}.embed"] = x
        for i in range(MIT_DEPTHS[s]):
            b = f"{p}block{s + 1}.{i}."
            y = F.layer_norm(x, (C,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], 1e-6)  # eps :519
            x = x + _attention(sd, b + "attn.", y, H, W, MIT_HEADS[s], MIT_SR[s])
            if taps is not None:
                taps[f"mit.s{s + 1}.b{i}.attn"] = x
            y = F.layer_norm(x, (C,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], 1e-6)
            x = x + _mix_ffn(sd, b + "mlp.", y, H, W)
            if taps is not None:
                taps[f"mit.s{s + 1}.b{i}"] = x
        x = F.layer_norm(x, (C,), sd[f"{p}norm{s + 1}.weight"], sd[f"{p}norm{s + 1}.bias"], 1e-6)
        if taps is not None:
            taps[f"mit.c{s + 1}"] = x  # tokens [B, H*W, C]
        x = x.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()
        outs.append(x)
    return outs


def low_level_encoder(sd, images):
    """perspectivefields.py:79-83: conv7x7/2 (no bias) -> BatchNorm2d (eval) -> ReLU."""
    x = F.conv2d(images, sd["ll_enc.conv1.weight"], None, stride=2, padding=3)
    x = F.batch_norm(x, sd["ll_enc.bn1.running_mean"], sd["ll_enc.bn1.running_var"], sd["ll_enc.bn1.weight"],
                     sd["ll_enc.bn1.bias"], False, 0.1, 1e-5)
    return F.relu(x)


# ------------------------------------------------------------------------------------------------ heads
def _rcu(sd, p, x):
    """decode_head.py:244-256.  ``self.relu`` is in place, so the skip adds relu(x), not x."""
    r = F.relu(x)
    out = F.conv2d(r, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    out = F.relu(out)
    out = F.conv2d(out, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    return out + r


def _fusion(sd, p, x0, x1=None):
    """decode_head.py:272-288 (FeatureFusionBlock.forward)."""
    out = x0
    if x1 is not None:
        out = out + _rcu(sd, p + "resConfUnit1.", x1)
    out = _rcu(sd, p + "resConfUnit2.", out)
    return F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=False)


def head_layers(sd, p, pred_name, hl, ll, taps=None, tag=""):
    """gravity_head.py:139-176 / latitude_head.py:138-175 (``layers``); MLP = decode_head.py:51-54."""
    c = hl
    fused = None
    for lvl in (4, 3, 2, 1):
        x = c[lvl - 1]
        n, _, h, w = x.shape
        t = F.linear(x.flatten(2).transpose(1, 2), sd[f"{p}linear_c{lvl}.proj.weight"], sd[f"{p}linear_c{lvl}.proj.bias"])
        t = t.permute(0, 2, 1).reshape(n, -1, h, w)
        t = F.conv2d(t, sd[f"{p}linear_c{lvl}_proc.weight"], sd[f"{p}linear_c{lvl}_proc.bias"], padding=1)
        if taps is not None:
            taps[f"{tag}.proc{lvl}"] = t
        fused = _fusion(sd, f"{p}fusion{lvl}.", t) if lvl == 4 else _fusion(sd, f"{p}fusion{lvl}.", fused, t)
        if taps is not None:
            taps[f"{tag}.fusion{lvl}"] = fused
    x = torch.cat([fused, ll], dim=1)
    x = F.relu(F.conv2d(x, sd[p + "conv_fuse_conv0.conv.weight"], sd[p + "conv_fuse_conv0.conv.bias"], padding=1))
    if taps is not None:
        taps[f"{tag}.conv0"] = x
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    x = F.relu(F.conv2d(x, sd[p + "conv_fuse_conv1.conv.weight"], sd[p + "conv_fuse_conv1.conv.bias"], padding=1))
    if taps is not None:
        taps[f"{tag}.conv1"] = x
    return F.conv2d(x, sd[p + pred_name + ".weight"], sd[p + pred_name + ".bias"])


def heads_inference(sd, cfg, hl, ll, taps=None):
    """persformer_heads.py:73-81; GravityDecoder.inference gravity_head.py:190-197 (the interpolate with
    scale_factor=1 is an identity); LatitudeDecoder.inference latitude_head.py:189-193."""
    g = head_layers(sd, "persformer_heads.gravity_head.", "linear_pred_gravity", hl, ll, taps, "g")
    if taps is not None:
        taps["g.raw"] = g
    if cfg["gravity"] == "regression":
        g = F.normalize(g, dim=1)
    l = head_layers(sd, "persformer_heads.latitude_head.", "linear_pred_latitude", hl, ll, taps, "l")
    if taps is not None:
        taps["l.raw"] = l
    if cfg["latitude"] == "regression":
        l = torch.clamp(l, -1, 1)
    return g, l


# --------------------------------------------------------------------------------------- post-processing
def decode_bin(angle_bin, num_bin):
    """utils/utils.py:114-130."""
    angle = (angle_bin * (360 / (num_bin - 1)) - 180) / 180 * np.pi
    vec = torch.stack((torch.cos(angle), torch.sin(angle)), dim=0)
    vec[:, angle_bin == num_bin - 1] = 0
    return vec


def decode_bin_latitude(binmap, num_classes):
    """utils/utils.py:148-162."""
    bin_size = 180 / num_classes
    centers = torch.arange(-90, 90, bin_size) + bin_size / 2
    return centers[binmap]


def pf_postprocess(result, out_h, out_w):
    """utils/utils.py:483-507: crop to the network size, bilinear (no antialias) to (H, W)."""
    result = result[:, :NET_H, :NET_W].expand(1, -1, -1, -1)
    return F.interpolate(result, size=(out_h, out_w), mode="bilinear", align_corners=False)[0]


def postprocess_gravity(cfg, result, height, width):
    """gravity_head.py:237-261."""
    vec = result if cfg["gravity"] == "regression" else decode_bin(result.argmax(dim=0), cfg["gravity_classes"])
    scale = torch.tensor([[width / NET_W], [height / NET_H]]).unsqueeze(-1)
    vec = pf_postprocess(vec * scale, height, width)
    return F.normalize(vec, dim=0)


def postprocess_latitude(cfg, result, height, width):
    """latitude_head.py:195-219."""
    if cfg["latitude"] == "regression":
        lat = pf_postprocess(result, height, width)[0]
        return torch.rad2deg(torch.asin(lat))
    lat = decode_bin_latitude(result.argmax(dim=0), cfg["latitude_classes"]).unsqueeze(0)
    return pf_postprocess(lat, height, width)[0]


# ---------------------------------------------------------------------------------------------- ParamNet
def _ln_channels_first(x, w, b, eps=1e-6):
    """convnext.py:177-182."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def convnext_t(sd, x, taps=None):
    """convnext.py:140-152 (forward), :46-59 (Block.forward)."""
    p = "param_net.backbone."
    for s in range(4):
        d = f"{p}downsample_layers.{s}."
        if s == 0:
            x = F.conv2d(x, sd[d + "0.weight"], sd[d + "0.bias"], stride=4)
            x = _ln_channels_first(x, sd[d + "1.weight"], sd[d + "1.bias"])
        else:
            x = _ln_channels_first(x, sd[d + "0.weight"], sd[d + "0.bias"])
            x = F.conv2d(x, sd[d + "1.weight"], sd[d + "1.bias"], stride=2)
        C = CNX_DIMS[s]
        for j in range(CNX_DEPTHS[s]):
            b = f"{p}stages.{s}.{j}."
            y = F.conv2d(x, sd[b + "dwconv.weight"], sd[b + "dwconv.bias"], padding=3, groups=C)
            y = y.permute(0, 2, 3, 1)
            y = F.layer_norm(y, (C,), sd[b + "norm.weight"], sd[b + "norm.bias"], 1e-6)
            y = F.linear(y, sd[b + "pwconv1.weight"], sd[b + "pwconv1.bias"])
            y = F.gelu(y)
            y = F.linear(y, sd[b + "pwconv2.weight"], sd[b + "pwconv2.bias"])
            y = sd[b + "gamma"] * y
            x = x + y.permute(0, 3, 1, 2)
        if taps is not None:
            taps[f"cnx.s{s}"] = x
    x = F.layer_norm(x.mean([-2, -1]), (CNX_DIMS[3],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    return F.linear(x, sd[p + "head.weight"], sd[p + "head.bias"])


def general_vfov_to_focal(rel_cx, rel_cy, h, gvfov, degree):
    """utils/utils.py:47-91, array branch (scipy.optimize.fsolve from 1.5, then abs)."""
    import scipy.optimize

    def fun(focal, *args):
        h, d_cx, d_cy, target = args
        p_sqr = (focal / h) ** 2 + d_cx ** 2 + (d_cy + 0.5) ** 2
        q_sqr = (focal / h) ** 2 + d_cx ** 2 + (d_cy - 0.5) ** 2
        return (p_sqr + q_sqr - 1) / 2 / np.sqrt(p_sqr) / np.sqrt(q_sqr) - target

    if degree:
        gvfov = np.radians(gvfov)
    focal = scipy.optimize.fsolve(fun, np.ones(len(rel_cx)) * 1.5, args=(h, rel_cx, rel_cy, np.cos(gvfov)))
    return np.abs(focal)


def param_net(sd, cfg, pred_gravity, pred_latitude, taps=None):
    """param_network.py:46-69 (ParamNet, eval) and :193-221 (ParamNetConvNextRegress, eval), followed by the
    key completion of perspectivefields.py:261-267."""
    images = torch.cat((pred_gravity, pred_latitude), dim=1)
    if cfg["param_net"] == "ParamNet":
        x = convnext_t(sd, images, taps)
        assert not cfg["recover_pp"]
        param = {
            "pred_roll": x[:, 0] * 90.0,
            "pred_pitch": x[:, 1] * 90.0,
            "pred_vfov": x[:, 2] * 90.0,
            "pred_rel_focal": 1 / 2 / torch.tan(x[:, 2]),  # sic: tan of the normalised value
        }
    else:
        images = F.interpolate(images, (cfg["input_size"], cfg["input_size"]))  # nearest
        x = convnext_t(sd, images, taps)
        factors = {"roll": 90.0, "pitch": 90.0, "vfov": 90.0, "rel_focal": 1.0, "rel_cx": 1.0, "rel_cy": 1.0,
                   "general_vfov": 90.0}
        param = {"pred_" + k: x[:, i] * factors[k] for i, k in enumerate(cfg["predict_params"])}
        param["pred_rel_focal"] = torch.FloatTensor(
            general_vfov_to_focal(param["pred_rel_cx"].numpy(), param["pred_rel_cy"].numpy(), 1,
                                  param["pred_general_vfov"].numpy(), degree=True))
    if taps is not None:
        taps["cnx.out"] = x
    if "pred_general_vfov" not in param:
        param["pred_general_vfov"] = param["pred_vfov"]
    if "pred_rel_cx" not in param:
        param["pred_rel_cx"] = torch.zeros_like(param["pred_vfov"])
    if "pred_rel_cy" not in param:
        param["pred_rel_cy"] = torch.zeros_like(param["pred_vfov"])
    return param


# ------------------------------------------------------------------------------------------- public API
def preprocess(img_bgr):
    """perspectivefields.py:196-202: copy, (BGR kept), Pillow resize to 320x320, float32 CHW."""
    assert img_bgr.dtype == np.uint8 and img_bgr.ndim == 3 and img_bgr.shape[2] == 3
    image = resize_bilinear_u8(img_bgr, NET_H, NET_W)
    return torch.as_tensor(image.astype("float32").transpose(2, 0, 1))


def resize_float(img, new_h, new_w):
    """perspectivefields.py:47-66, the non-uint8 branch of ResizeTransform.apply_image: HW(C) numpy -> NCHW torch ->
    ``F.interpolate(mode="bilinear", align_corners=False)`` (no antialias) -> HW(C) numpy of the input dtype."""
    if any(x < 0 for x in img.strides):
        img = np.ascontiguousarray(img)
    t = torch.from_numpy(img)
    shape = list(t.shape)
    shape_4d = shape[:2] + [1] * (4 - len(shape)) + shape[2:]
    t = t.view(shape_4d).permute(2, 3, 0, 1)
    t = F.interpolate(t, (new_h, new_w), mode="bilinear", align_corners=False)
    shape[:2] = (new_h, new_w)
    return t.permute(2, 3, 0, 1).view(shape).numpy()


def inference_float(sd, version, img_bgr):
    """perspectivefields.py:194-205 for a non-uint8 image: float resize branch, then ``astype("float32")`` and forward."""
    image = resize_float(img_bgr.copy(), NET_H, NET_W)
    image = torch.as_tensor(image.astype("float32").transpose(2, 0, 1))
    return forward(sd, version, [{"image": image, "height": img_bgr.shape[0], "width": img_bgr.shape[1]}])[0]


@torch.no_grad()
def forward(sd, version, batched_inputs, taps=None):
    """perspectivefields.py:223-272 on CPU fp32.  ``batched_inputs``: list of {"image","height","width"}."""
    cfg = VARIANTS[version]
    mean = torch.tensor(PIXEL_MEAN).view(-1, 1, 1)
    std = torch.tensor(PIXEL_STD).view(-1, 1, 1)
    images = torch.stack([(x["image"] - mean) / std for x in batched_inputs])
    hl = mit_b3(sd, images, taps)
    ll = low_level_encoder(sd, images)
    if taps is not None:
        taps["ll"] = ll
    g, l = heads_inference(sd, cfg, hl, ll, taps)
    results = []
    for i, inp in enumerate(batched_inputs):
        h, w = inp["height"], inp["width"]
        results.append({
            "pred_gravity": g[i],
            "pred_gravity_original": postprocess_gravity(cfg, g[i], h, w),
            "pred_latitude": l[i],
            "pred_latitude_original": postprocess_latitude(cfg, l[i], h, w),
            "pred_latitude_original_mode": "deg",
        })
    if cfg["param_net"] is not None:
        param = param_net(sd, cfg, g, l, taps)
        for i in range(len(results)):
            results[i].update({k: v[i] for k, v in param.items()})
    return results


def inference_batch(sd, version, img_bgr_list, taps=None):
    """perspectivefields.py:207-221."""
    inputs = [{"image": preprocess(im), "height": im.shape[0], "width": im.shape[1]} for im in img_bgr_list]
    return forward(sd, version, inputs, taps)


def inference(sd, version, img_bgr):
    """perspectivefields.py:194-205."""
    return inference_batch(sd, version, [img_bgr])[0]
