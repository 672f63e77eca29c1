"""GPU: every operator of libpf_b200.so, called through the C ABI, against a float64 torch restatement of the same op
on the same seeded inputs.  Tolerances: 5e-5 relative for the bf16x3 split-precision GEMM engine (per-product error
~2^-17; the end-to-end bar is 1e-3), 1e-5 for fp32 CUDA-core ops, bit-exact for the integer resize."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import pf_test_util as U

pytestmark = pytest.mark.gpu


def _rn(g, *s):
    return torch.randn(*s, generator=g)


# B, H, W, Cin, N, K, stride, pad, in_relu, act, res, res_relu
GEMM_CASES = [
    (2, 20, 20, 64, 256, 3, 1, 1, 0, 0, 0, 0),      # composed head conv shape (C1 -> 256)
    (2, 20, 20, 256, 256, 3, 1, 1, 1, 1, 0, 0),     # RCU conv1: relu prologue + relu epilogue
    (1, 23, 17, 256, 256, 3, 1, 1, 0, 0, 1, 1),     # RCU conv2: + relu(residual), ragged M
    (1, 1, 700, 320, 640, 1, 1, 0, 0, 0, 0, 0),     # kv linear (N tail: 640 = 5 x 128)
    (1, 1, 300, 96, 384, 1, 1, 0, 0, 2, 0, 0),      # ConvNeXt pwconv1 + GELU
    (1, 1, 130, 384, 96, 1, 1, 0, 0, 0, 1, 0),      # pwconv2 + residual, N = 96 (64-wide tiles)
    (2, 80, 80, 64, 64, 8, 8, 0, 0, 0, 0, 0),       # spatial-reduction conv k = s = 8
    (2, 40, 40, 64, 128, 3, 2, 1, 0, 0, 0, 0),      # overlap patch embed, stride 2
    (1, 24, 24, 64, 32, 3, 1, 1, 0, 1, 0, 0),       # conv_fuse_conv1, N = 32 tile
    (1, 16, 16, 96, 192, 2, 2, 0, 0, 0, 0, 0),      # ConvNeXt downsample
    (1, 40, 40, 320, 64, 3, 1, 1, 0, 1, 0, 0),      # conv_fuse_conv0 shape
    (1, 1, 1, 512, 512, 1, 1, 0, 0, 0, 0, 0),       # single row (M = 1)
]


@pytest.mark.parametrize("case", GEMM_CASES)
def test_conv_gemm(case):
    B, H, W, Cin, N, K, s, p, ir, act, res, rr = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = _rn(g, B, Cin, H, W).cuda()
    w = _rn(g, N, Cin, K, K) / (Cin * K * K) ** 0.5
    b = _rn(g, N)
    ref = F.conv2d((F.relu(x) if ir else x).double(), w.double().cuda(), b.double().cuda(), stride=s, padding=p)
    ref = F.relu(ref) if act == 1 else (F.gelu(ref) if act == 2 else ref)
    r = None
    if res:
        r = _rn(g, *ref.shape).cuda()
        ref = ref + (F.relu(r) if rr else r).double()
        r = r.permute(0, 2, 3, 1).contiguous()
    y = U.conv_gemm(x.permute(0, 2, 3, 1).contiguous(), w, b, s, p, ir, act, r, rr)
    assert U.rel_err(y.permute(0, 3, 1, 2), ref) < 5e-5


def test_conv_gemm_is_deterministic_and_batch_invariant():
    g = torch.Generator().manual_seed(3)
    x = _rn(g, 3, 12, 12, 256).cuda()
    w, b = _rn(g, 256, 256, 3, 3) / 48, _rn(g, 256)
    y = U.conv_gemm(x, w, b, 1, 1)
    assert torch.equal(y, U.conv_gemm(x, w, b, 1, 1))
    assert torch.equal(y[1:2], U.conv_gemm(x[1:2].contiguous(), w, b, 1, 1))


@pytest.mark.parametrize("C", [64, 96, 128, 192, 320, 384, 512, 768])
def test_layernorm(C):
    from perspectivefields_b200 import _native

    g = torch.Generator().manual_seed(C)
    x = (_rn(g, 777, C) * 3 + 1).cuda()
    w, b = _rn(g, C).cuda(), _rn(g, C).cuda()
    y = torch.empty_like(x)
    _native.check(_native.lib().pf_op_layernorm(x.data_ptr(), y.data_ptr(), 777, C, w.data_ptr(), b.data_ptr(), 1e-6, U.stream_ptr()))
    assert U.rel_err(y, F.layer_norm(x.double(), (C,), w.double(), b.double(), 1e-6)) < 1e-5


@pytest.mark.parametrize("engine", ["fp32", "mma"])
@pytest.mark.parametrize("shape", [(2, 6400, 1), (2, 1600, 2), (1, 400, 5), (1, 100, 8), (1, 77, 2), (1, 1000, 1)])
def test_attention(shape, engine):
    from perspectivefields_b200 import _native

    B, N, heads = shape
    C = heads * 64
    g = torch.Generator().manual_seed(N)
    q, kv = (_rn(g, B, N, C) * 2).cuda(), (_rn(g, B, 100, 2 * C) * 2).cuda()
    o = torch.empty_like(q)
    fn = _native.lib().pf_op_attention if engine == "fp32" else _native.lib().pf_op_attention_mma
    _native.check(fn(q.data_ptr(), kv.data_ptr(), o.data_ptr(), B, N, C, heads, U.stream_ptr()))
    qh = q.double().reshape(B, N, heads, 64).permute(0, 2, 1, 3)
    kvh = kv.double().reshape(B, 100, 2, heads, 64).permute(2, 0, 3, 1, 4)
    ref = ((qh @ kvh[0].transpose(-2, -1)) * 0.125).softmax(-1) @ kvh[1]
    assert U.rel_err(o, ref.transpose(1, 2).reshape(B, N, C)) < (1e-5 if engine == "fp32" else 5e-5)


def test_depthwise_and_upsample():
    from perspectivefields_b200 import _native

    L = _native.lib()
    g = torch.Generator().manual_seed(9)
    x, w, b = _rn(g, 2, 256, 20, 20).cuda(), _rn(g, 256, 1, 3, 3).cuda(), _rn(g, 256).cuda()
    xh = x.permute(0, 2, 3, 1).contiguous()
    y = torch.empty_like(xh)
    _native.check(L.pf_op_dwconv3x3_gelu(xh.data_ptr(), y.data_ptr(), 2, 20, 20, 256, w.reshape(256, 9).t().contiguous().data_ptr(), b.data_ptr(), U.stream_ptr()))
    assert U.rel_err(y.permute(0, 3, 1, 2), F.gelu(F.conv2d(x.double(), w.double(), b.double(), padding=1, groups=256))) < 1e-5
    x, w, b = _rn(g, 2, 96, 16, 16).cuda(), _rn(g, 96, 1, 7, 7).cuda(), _rn(g, 96).cuda()
    xh = x.permute(0, 2, 3, 1).contiguous()
    y = torch.empty_like(xh)
    _native.check(L.pf_op_dwconv7x7(xh.data_ptr(), y.data_ptr(), 2, 16, 16, 96, w.reshape(96, 49).t().contiguous().data_ptr(), b.data_ptr(), U.stream_ptr()))
    assert U.rel_err(y.permute(0, 3, 1, 2), F.conv2d(x.double(), w.double(), b.double(), padding=3, groups=96)) < 1e-5
    x = _rn(g, 2, 64, 10, 13).cuda()
    xh = x.permute(0, 2, 3, 1).contiguous()
    y = torch.empty(2, 20, 26, 64, device="cuda")
    _native.check(L.pf_op_upsample2x(xh.data_ptr(), y.data_ptr(), 2, 10, 13, 64, U.stream_ptr()))
    assert U.rel_err(y.permute(0, 3, 1, 2), F.interpolate(x.double(), scale_factor=2, mode="bilinear", align_corners=False)) < 1e-6


@pytest.mark.parametrize("hw", [(480, 640), (240, 320), (320, 320), (721, 900), (1536, 2048), (33, 47), (512, 512), (320, 500), (1, 1)])
def test_preprocess_is_bit_exact_vs_pillow(hw):
    from PIL import Image

    from perspectivefields_b200 import _native

    h, w = hw
    img = np.random.RandomState(h + w).randint(0, 256, (h, w, 3), dtype=np.uint8)
    mean = np.array([103.53, 116.28, 123.675], np.float32)
    ref = np.asarray(Image.fromarray(img).resize((320, 320), Image.BILINEAR)).astype(np.float32) - mean
    d = torch.from_numpy(img).cuda()
    y = torch.empty(320, 320, 4, device="cuda")
    _native.check(_native.lib().pf_op_preprocess(d.data_ptr(), h, w, (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(1, 1, 1), y.data_ptr(), U.stream_ptr()))
    torch.cuda.synchronize()
    got = y.cpu().numpy()
    assert np.array_equal(got[..., :3], ref) and not got[..., 3].any()


# ---- tcgen05 / TMEM engine: same math, 128 x {256,128,64,32} tiles
TC_CASES = [
    (2, 20, 20, 256, 256, 3, 1, 1, 1, 1, 0, 0),     # RCU conv1
    (1, 23, 17, 256, 256, 3, 1, 1, 0, 0, 1, 1),     # RCU conv2 + relu(residual), ragged M (391 rows)
    (2, 16, 16, 64, 512, 3, 1, 1, 0, 0, 0, 0),      # composed proc conv (two N tiles)
    (1, 1, 700, 320, 256, 1, 1, 0, 0, 0, 0, 0),     # plain linear, K = 320 (10 k-steps, ring wraps)
    (1, 1, 64, 32, 256, 1, 1, 0, 0, 0, 0, 0),       # single k-step
    (3, 40, 40, 256, 256, 3, 1, 1, 1, 0, 0, 0),     # 38 tiles, 72 k-steps
    (1, 1, 700, 320, 640, 1, 1, 0, 0, 0, 0, 0),     # N tile 128 (kv linear)
    (1, 1, 300, 96, 384, 1, 1, 0, 0, 2, 0, 0),      # N tile 128 + GELU
    (1, 1, 130, 384, 96, 1, 1, 0, 0, 0, 1, 0),      # N tile 32 (x3) + residual
    (2, 80, 80, 64, 64, 8, 8, 0, 0, 0, 0, 0),       # N tile 64, k = s = 8
    (1, 24, 24, 64, 32, 3, 1, 1, 0, 1, 0, 0),       # N tile 32
    (1, 40, 40, 320, 64, 3, 1, 1, 0, 1, 0, 0),      # N tile 64, K = 2880
    (2, 40, 40, 64, 128, 3, 2, 1, 0, 0, 0, 0),      # N tile 128, stride 2
    (1, 1, 500, 64, 320, 1, 1, 0, 0, 0, 0, 0),      # N = 320 -> five 64-wide tiles
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_gemm_tcgen05(case):
    B, H, W, Cin, N, K, s, p, ir, act, res, rr = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = _rn(g, B, Cin, H, W).cuda()
    w = _rn(g, N, Cin, K, K) / (Cin * K * K) ** 0.5
    b = _rn(g, N)
    ref = F.conv2d((F.relu(x) if ir else x).double(), w.double().cuda(), b.double().cuda(), stride=s, padding=p)
    ref = F.relu(ref) if act == 1 else (F.gelu(ref) if act == 2 else ref)
    r = None
    if res:
        r = _rn(g, *ref.shape).cuda()
        ref = ref + (F.relu(r) if rr else r).double()
        r = r.permute(0, 2, 3, 1).contiguous()
    xh = x.permute(0, 2, 3, 1).contiguous()
    y = U.conv_gemm(xh, w, b, s, p, ir, act, r, rr, engine=1)
    assert U.rel_err(y.permute(0, 3, 1, 2), ref) < 5e-5
    # the two engines evaluate the same three bf16 products per term; only the fp32 summation order differs
    y0 = U.conv_gemm(xh, w, b, s, p, ir, act, r, rr, engine=0)
    assert U.rel_err(y, y0) < 2e-6


# ---- halo-tile tcgen05 kernel for 3x3 / stride 1 / pad 1 convolutions (16 x 8 pixel tiles, partial tiles masked)
HALO_CASES = [
    (2, 16, 8, 64, 32, 0, 1, 0, 0),       # exactly one tile per image, conv_fuse_conv1 shape class
    (1, 80, 80, 256, 256, 1, 1, 0, 0),    # RCU conv1 at 80x80
    (2, 23, 17, 256, 256, 0, 0, 1, 1),    # ragged tiles in both directions + relu(residual)
    (1, 10, 10, 512, 512, 0, 0, 0, 0),    # proc conv at the coarsest level (tile larger than the image), two N tiles
    (1, 40, 40, 320, 64, 0, 1, 0, 0),     # conv_fuse_conv0 shape class (5 chunks)
    (3, 33, 9, 128, 128, 0, 0, 0, 0),
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv3x3_halo_tcgen05(case):
    B, H, W, Cin, N, ir, act, res, rr = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = _rn(g, B, Cin, H, W).cuda()
    w = _rn(g, N, Cin, 3, 3) / (Cin * 9) ** 0.5
    b = _rn(g, N)
    ref = F.conv2d((F.relu(x) if ir else x).double(), w.double().cuda(), b.double().cuda(), padding=1)
    ref = F.relu(ref) if act == 1 else ref
    r = None
    if res:
        r = _rn(g, *ref.shape).cuda()
        ref = ref + (F.relu(r) if rr else r).double()
        r = r.permute(0, 2, 3, 1).contiguous()
    xh = x.permute(0, 2, 3, 1).contiguous()
    y = U.conv_gemm(xh, w, b, 1, 1, ir, act, r, rr, engine=2)
    assert U.rel_err(y.permute(0, 3, 1, 2), ref) < 5e-5
    # same products, different fp32 summation order (channel chunks outermost instead of filter taps)
    assert U.rel_err(y, U.conv_gemm(xh, w, b, 1, 1, ir, act, r, rr, engine=0)) < 2e-5


# ---- TMA -> tcgen05 engine (engine 3): inputs are split into bf16 hi/lo planes first, exactly as the forward graph does
@pytest.mark.parametrize("case", GEMM_CASES + TC_CASES[:3])
def test_conv_gemm_tma_engine(case):
    B, H, W, Cin, N, K, s, p, ir, act, res, rr = case
    if N % 32:
        pytest.skip("TMA engine needs N % 32 == 0")
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = _rn(g, B, Cin, H, W).cuda()
    w = _rn(g, N, Cin, K, K) / (Cin * K * K) ** 0.5
    b = _rn(g, N)
    ref = F.conv2d((F.relu(x) if ir else x).double(), w.double().cuda(), b.double().cuda(), stride=s, padding=p)
    ref = F.relu(ref) if act == 1 else (F.gelu(ref) if act == 2 else ref)
    r = None
    if res:
        r = _rn(g, *ref.shape).cuda()
        ref = ref + (F.relu(r) if rr else r).double()
        r = r.permute(0, 2, 3, 1).contiguous()
    y = U.conv_gemm(x.permute(0, 2, 3, 1).contiguous(), w, b, s, p, ir, act, r, rr, engine=3)
    assert U.rel_err(y.permute(0, 3, 1, 2), ref) < 5e-5


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv3x3_tma_halo(case):
    B, H, W, Cin, N, ir, act, res, rr = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = _rn(g, B, Cin, H, W).cuda()
    w = _rn(g, N, Cin, 3, 3) / (Cin * 9) ** 0.5
    b = _rn(g, N)
    ref = F.conv2d((F.relu(x) if ir else x).double(), w.double().cuda(), b.double().cuda(), padding=1)
    ref = F.relu(ref) if act == 1 else ref
    r = None
    if res:
        r = _rn(g, *ref.shape).cuda()
        ref = ref + (F.relu(r) if rr else r).double()
        r = r.permute(0, 2, 3, 1).contiguous()
    y = U.conv_gemm(x.permute(0, 2, 3, 1).contiguous(), w, b, 1, 1, ir, act, r, rr, engine=3)
    assert U.rel_err(y.permute(0, 3, 1, 2), ref) < 5e-5
