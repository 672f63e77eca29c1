"""GPU: time (and let ncu capture) single GEMM-engine launches of chosen shapes through pf_op_conv_gemm."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pf_test_util as U
from perspectivefields_b200 import _native
L = _native.lib()
g = torch.Generator().manual_seed(0)
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(12800, 1280, 320), (12800, 320, 320), (12800, 320, 1280), (204800, 384, 96), (204800, 64, 64)]
for (M, N, K) in shapes:
    x = torch.randn(1, 1, M, K, generator=g).cuda()
    w = torch.randn(N, K, 1, 1, generator=g) / K ** 0.5
    hi, lo = U.split_hi_lo(w.reshape(N, K)); hi, lo = hi.cuda(), lo.cuda()
    b = torch.randn(N, generator=g).cuda()
    y = torch.empty(1, 1, M, N, device="cuda")
    ACT = int(os.environ.get('PF_PROBE_ACT', '0'))
    for eng in (3,):
        for _ in range(2):
            _native.check(L.pf_op_conv_gemm(x.data_ptr(), 1, 1, M, K, hi.data_ptr(), lo.data_ptr(), b.data_ptr(), N, 1, 1, 1, 0, 0, ACT, None, 0, y.data_ptr(), U.stream_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the test entry splits the input and synchronises inside: time includes the split kernel; ncu isolates the GEMM
        e0.record()
        for _ in range(5):
            _native.check(L.pf_op_conv_gemm(x.data_ptr(), 1, 1, M, K, hi.data_ptr(), lo.data_ptr(), b.data_ptr(), N, 1, 1, 1, 0, 0, ACT, None, 0, y.data_ptr(), U.stream_ptr()))
        e1.record(); torch.cuda.synchronize()
        print(f"M{M} N{N} K{K} engine {eng}: {e0.elapsed_time(e1)/5*1000:.1f} us per call (incl. split + sync)", flush=True)
