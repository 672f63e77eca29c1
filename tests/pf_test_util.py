"""Shared helpers for the GPU parity tests: synthetic checkpoints in a private hub cache, op wrappers over the C ABI."""
import ctypes
import os
import tempfile

import numpy as np
import torch

_HUB = None


def hub_dir():
    """Private TORCH_HOME so that the product's normal hub-cache loader finds the synthetic checkpoints."""
    global _HUB
    if _HUB is None:
        _HUB = tempfile.mkdtemp(prefix="pf_hub_")
        os.environ["TORCH_HOME"] = _HUB
        os.makedirs(os.path.join(_HUB, "hub", "checkpoints"), exist_ok=True)
    return os.path.join(_HUB, "hub", "checkpoints")


def write_synthetic_checkpoint(version, seed=0, **kw):
    from oracle import weights_gen as wg
    from oracle.variants import VARIANTS

    sd = wg.synth_state_dict(version, seed, **kw)
    torch.save({"model": sd}, os.path.join(hub_dir(), VARIANTS[version]["ckpt"]))
    return sd


def make_model(version, seed=0, device="cuda", model_kwargs=None, **kw):
    """(product model on `device`, reference-layout state dict it was loaded from)."""
    sd = write_synthetic_checkpoint(version, seed, **kw)
    from perspectivefields_b200 import PerspectiveFields

    m = PerspectiveFields(version, **(model_kwargs or {})).eval()
    if device is not None:
        m = m.to(device)
    return m, sd


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def split_hi_lo(w):
    from perspectivefields_b200.weights import split_hi_lo as s

    return s(w)


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def conv_gemm(x_nhwc, w_oihw, bias, stride, pad, in_relu=False, act=0, res=None, res_relu=False):
    """Run pf_op_conv_gemm.  x: [B,H,W,Cin] cuda fp32; w: [N,Cin,KH,KW] (cpu or cuda)."""
    from perspectivefields_b200 import _native

    L = _native.lib()
    B, H, W, Cin = x_nhwc.shape
    N, _, KH, KW = w_oihw.shape
    wnk = w_oihw.permute(0, 2, 3, 1).reshape(N, -1).cpu()
    hi, lo = split_hi_lo(wnk)
    hi, lo = hi.cuda(), lo.cuda()
    OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    y = torch.empty((B, OH, OW, N), dtype=torch.float32, device="cuda")
    b = bias.float().cuda().contiguous() if bias is not None else None
    r = res.contiguous() if res is not None else None
    _native.check(L.pf_op_conv_gemm(x_nhwc.contiguous().data_ptr(), B, H, W, Cin, hi.data_ptr(), lo.data_ptr(),
                                    b.data_ptr() if b is not None else None, N, KH, KW, stride, pad, int(in_relu), act,
                                    r.data_ptr() if r is not None else None, int(res_relu), y.data_ptr(), stream_ptr()))
    torch.cuda.synchronize()
    return y


def stable_mask(ref_logits, err, out_h=None, out_w=None, factor=4.0):
    """Pixels of a decoded (argmax) field that do not depend on a near-tie of the logits.  ``ref_logits``: [NC, 320, 320] oracle
    logits; a 320x320 pixel is stable when its top-2 margin exceeds ``factor`` x the logit error ``err``.  For a field resampled to
    (out_h, out_w) a pixel is stable when all four bilinear source taps are: the unstable mask is dilated by one pixel and
    resampled with the same bilinear map (any contribution of an unstable pixel marks the output pixel)."""
    import torch.nn.functional as F

    top2 = ref_logits.topk(2, dim=0).values
    unstable = ((top2[0] - top2[1]) <= factor * err).float()[None, None]
    if out_h is None:
        return unstable[0, 0] == 0
    unstable = F.max_pool2d(unstable, 3, stride=1, padding=1)
    up = F.interpolate(unstable, size=(out_h, out_w), mode="bilinear", align_corners=False)
    return up[0, 0] == 0
