"""GPU: first contact with the tcgen05 kernel -- prints relative errors, never asserts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import pf_test_util as U
g = torch.Generator().manual_seed(1)
rn = lambda *s: torch.randn(*s, generator=g)
cases = [(1, 1, 128, 32, 256, 1, 1, 0, 0, 0, 0, 0), (1, 1, 128, 64, 256, 1, 1, 0, 0, 0, 0, 0), (1, 1, 300, 320, 256, 1, 1, 0, 0, 0, 0, 0),
         (2, 20, 20, 256, 256, 3, 1, 1, 1, 1, 0, 0), (1, 23, 17, 256, 256, 3, 1, 1, 0, 0, 1, 1), (2, 16, 16, 64, 512, 3, 1, 1, 0, 0, 0, 0),
         (3, 40, 40, 256, 256, 3, 1, 1, 1, 0, 0, 0), (1, 1, 700, 320, 640, 1, 1, 0, 0, 0, 0, 0), (1, 1, 130, 384, 96, 1, 1, 0, 0, 0, 1, 0),
         (2, 80, 80, 64, 64, 8, 8, 0, 0, 0, 0, 0), (1, 24, 24, 64, 32, 3, 1, 1, 0, 1, 0, 0), (1, 1, 500, 64, 320, 1, 1, 0, 0, 0, 0, 0)]
for (B, H, W, Cin, N, K, s, p, ir, act, res, rr) in cases:
    x = rn(B, Cin, H, W).cuda(); w = rn(N, Cin, K, K) / (Cin * K * K) ** 0.5; b = rn(N)
    ref = F.conv2d((F.relu(x) if ir else x).double(), w.double().cuda(), b.double().cuda(), stride=s, padding=p)
    if act == 1: ref = F.relu(ref)
    r = None
    if res:
        r = rn(*ref.shape).cuda(); ref = ref + (F.relu(r) if rr else r).double(); r = r.permute(0, 2, 3, 1).contiguous()
    xh = x.permute(0, 2, 3, 1).contiguous()
    try:
        y = U.conv_gemm(xh, w, b, s, p, ir, act, r, rr, engine=1)
        e = U.rel_err(y.permute(0, 3, 1, 2), ref)
        y0 = U.conv_gemm(xh, w, b, s, p, ir, act, r, rr, engine=0)
        print(f"tc B{B} {H}x{W} Cin{Cin} N{N} k{K}: rel vs fp64 {e:.3g}; vs HMMA {U.rel_err(y, y0):.3g}", flush=True)
        if e > 1e-3:
            d = (y.permute(0, 3, 1, 2).double() - ref).abs()
            print("   bad: max at", [int(v) for v in torch.unravel_index(d.argmax(), d.shape)], "col-err profile", d.amax(dim=(0, 2, 3))[:8].tolist(),
                  "row-err", d.amax(dim=(0, 1)).flatten()[:8].tolist(), flush=True)
    except Exception as ex:
        print("EXC", type(ex).__name__, str(ex)[:300], flush=True)
        break
print("--- TMA engine (engine 3)", flush=True)
for (B, H, W, Cin, N, K, s, p, ir, act, res, rr) in [(1, 1, 128, 32, 256, 1, 1, 0, 0, 0, 0, 0), (1, 1, 300, 320, 256, 1, 1, 0, 0, 0, 0, 0), (1, 1, 700, 320, 640, 1, 1, 0, 0, 0, 0, 0),
        (1, 1, 130, 384, 96, 1, 1, 0, 0, 0, 1, 0), (1, 1, 5000, 64, 320, 1, 1, 0, 0, 2, 0, 0), (2, 80, 80, 64, 64, 8, 8, 0, 0, 0, 0, 0), (2, 40, 40, 64, 128, 3, 2, 1, 0, 0, 0, 0),
        (1, 16, 8, 64, 32, 3, 1, 1, 0, 0, 0, 0), (2, 23, 17, 256, 256, 3, 1, 1, 1, 1, 1, 1), (1, 10, 10, 512, 512, 3, 1, 1, 0, 0, 0, 0), (1, 40, 40, 320, 64, 3, 1, 1, 0, 1, 0, 0), (4, 80, 80, 256, 256, 3, 1, 1, 1, 1, 0, 0)]:
    x = rn(B, Cin, H, W).cuda(); w = rn(N, Cin, K, K) / (Cin * K * K) ** 0.5; b = rn(N)
    ref = F.conv2d((F.relu(x) if ir else x).double(), w.double().cuda(), b.double().cuda(), stride=s, padding=p)
    if act == 1: ref = F.relu(ref)
    if act == 2: ref = F.gelu(ref)
    r = None
    if res:
        r = rn(*ref.shape).cuda(); ref = ref + (F.relu(r) if rr else r).double(); r = r.permute(0, 2, 3, 1).contiguous()
    xh = x.permute(0, 2, 3, 1).contiguous()
    try:
        y = U.conv_gemm(xh, w, b, s, p, ir, act, r, rr, engine=3)
        e = U.rel_err(y.permute(0, 3, 1, 2), ref)
        print(f"tma B{B} {H}x{W} Cin{Cin} N{N} k{K} s{s}: rel vs fp64 {e:.3g}", flush=True)
        if e > 1e-3:
            d = (y.permute(0, 3, 1, 2).double() - ref).abs()
            print("   bad: argmax", [int(v) for v in torch.unravel_index(d.argmax(), d.shape)], "per-chan max", [round(float(v), 3) for v in d.amax(dim=(0, 2, 3))[:12]],
                  "per-row max", [round(float(v), 3) for v in d.amax(dim=(0, 1, 3))[:20]], flush=True)
    except Exception as ex:
        print("EXC", type(ex).__name__, str(ex)[:300], flush=True)
        break
print("--- halo-tile 3x3 kernel", flush=True)
for (B, H, W, Cin, N, ir, act, res, rr) in [(1, 16, 8, 64, 32, 0, 0, 0, 0), (2, 16, 8, 64, 32, 0, 1, 0, 0), (1, 80, 80, 256, 256, 1, 1, 0, 0), (2, 23, 17, 256, 256, 0, 0, 1, 1),
                                             (1, 10, 10, 512, 512, 0, 0, 0, 0), (1, 40, 40, 320, 64, 0, 1, 0, 0), (3, 33, 9, 128, 128, 0, 0, 0, 0)]:
    x = rn(B, Cin, H, W).cuda(); w = rn(N, Cin, 3, 3) / (Cin * 9) ** 0.5; b = rn(N)
    ref = F.conv2d((F.relu(x) if ir else x).double(), w.double().cuda(), b.double().cuda(), padding=1)
    if act == 1: ref = F.relu(ref)
    r = None
    if res:
        r = rn(*ref.shape).cuda(); ref = ref + (F.relu(r) if rr else r).double(); r = r.permute(0, 2, 3, 1).contiguous()
    xh = x.permute(0, 2, 3, 1).contiguous()
    try:
        y = U.conv_gemm(xh, w, b, 1, 1, ir, act, r, rr, engine=2)
        e = U.rel_err(y.permute(0, 3, 1, 2), ref)
        print(f"halo B{B} {H}x{W} Cin{Cin} N{N}: rel vs fp64 {e:.3g}", flush=True)
        if e > 1e-3:
            d = (y.permute(0, 3, 1, 2).double() - ref).abs()
            print("   bad: argmax", [int(v) for v in torch.unravel_index(d.argmax(), d.shape)], "per-row max", [round(float(v), 3) for v in d.amax(dim=(0, 1, 3))[:18]],
                  "per-col max", [round(float(v), 3) for v in d.amax(dim=(0, 1, 2))[:10]], flush=True)
    except Exception as ex:
        print("EXC", type(ex).__name__, str(ex)[:300], flush=True)
        break
# timing of the dominant shape: 3x3 256->256 at 80x80, batch 8, two groups emulated by N=256 launches
B, H, W, Cin, N = 8, 80, 80, 256, 256
x = rn(B, H, W, Cin).cuda(); w = rn(N, Cin, 3, 3) / 48; b = rn(N)
for eng in (0, 1, 2):
    try:
        for _ in range(2): U.conv_gemm(x, w, b, 1, 1, engine=eng)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        from perspectivefields_b200 import _native
        L = _native.lib()
        hi, lo = U.split_hi_lo(w.permute(0, 2, 3, 1).reshape(N, -1)); hi, lo = hi.cuda(), lo.cuda(); bb = b.cuda(); y = torch.empty(B, H, W, N, device="cuda")
        e0.record()
        for _ in range(10):
            L.pf_op_conv_gemm(x.data_ptr(), B, H, W, Cin, hi.data_ptr(), lo.data_ptr(), bb.data_ptr(), N, 3, 3, 1, 1, 0, 0, None, 0, y.data_ptr(), eng, U.stream_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"engine {eng}: {ms:.3f} ms -> {2*B*H*W*N*Cin*9/ms/1e9:.1f} TFLOP/s algorithmic", flush=True)
        if eng == 2:   # the narrow layers the halo kernel is for
            for (b2, h2, cin2, n2) in ((4, 160, 320, 64), (4, 320, 64, 32)):
                x2 = rn(b2, h2, h2, cin2).cuda(); w2 = rn(n2, cin2, 3, 3) / 50; hi2, lo2 = U.split_hi_lo(w2.permute(0, 2, 3, 1).reshape(n2, -1)); hi2, lo2 = hi2.cuda(), lo2.cuda()
                y2 = torch.empty(b2, h2, h2, n2, device="cuda"); bb2 = rn(n2).cuda()
                for e_ in (1, 2):
                    for _ in range(2): L.pf_op_conv_gemm(x2.data_ptr(), b2, h2, h2, cin2, hi2.data_ptr(), lo2.data_ptr(), bb2.data_ptr(), n2, 3, 3, 1, 1, 0, 1, None, 0, y2.data_ptr(), e_, U.stream_ptr())
                    e0.record()
                    for _ in range(5): L.pf_op_conv_gemm(x2.data_ptr(), b2, h2, h2, cin2, hi2.data_ptr(), lo2.data_ptr(), bb2.data_ptr(), n2, 3, 3, 1, 1, 0, 1, None, 0, y2.data_ptr(), e_, U.stream_ptr())
                    e1.record(); torch.cuda.synchronize()
                    ms2 = e0.elapsed_time(e1) / 5
                    print(f"  {h2}x{h2} Cin{cin2} N{n2} B{b2} engine {e_}: {ms2:.3f} ms -> {2*b2*h2*h2*n2*cin2*9/ms2/1e9:.1f} TFLOP/s", flush=True)
    except Exception as ex:
        print("EXC timing", eng, str(ex)[:300], flush=True)
