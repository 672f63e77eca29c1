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


def write_synthetic_checkpoint(version, seed=0):
    from oracle import weights_gen as wg
    from oracle.variants import VARIANTS

    sd = wg.synth_state_dict(version, seed)
    torch.save({"model": sd}, os.path.join(hub_dir(), VARIANTS[version]["ckpt"]))
    return sd


def make_model(version, seed=0, device="cuda"):
    """(product model on `device`, reference-layout state dict it was loaded from)."""
    sd = write_synthetic_checkpoint(version, seed)
    from perspectivefields_b200 import PerspectiveFields

    m = PerspectiveFields(version).eval()
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


def conv_gemm(x_nhwc, w_oihw, bias, stride, pad, in_relu=False, act=0, res=None, res_relu=False, engine=0):
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
                                    r.data_ptr() if r is not None else None, int(res_relu), y.data_ptr(), engine, stream_ptr()))
    torch.cuda.synchronize()
    return y
