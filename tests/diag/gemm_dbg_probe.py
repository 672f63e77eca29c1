"""GPU: where does a GEMM-mode launch spend its time?  Runs pf_op_conv_gemm for a few (M, N, K) shapes with the kernel's timing
switches (TmaGemmParams::dbg, env PF_GEMM_DBG) and, when run under
    ncu --metrics gpu__time_duration.sum -k regex:gemm --csv --log-file gpurun_out/gemm_dbg.csv python tests/diag/gemm_dbg_probe.py
gives the exact duration of every variant (launch order = the order printed here).  PF_NO_PAIR=1 selects the single-CTA kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pf_test_util as U
from perspectivefields_b200 import _native
L = _native.lib()
g = torch.Generator().manual_seed(0)
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(12800, 1280, 320), (12800, 320, 1280), (12800, 320, 320)]
DBG = [int(v) for v in os.environ.get('PF_DBG_LIST', '0,1,8,9,2,4,6,15').split(',')]   # all / no stores / no tmem read / neither / no loads / no MMA / no loads+no MMA / nothing but barriers
for (M, N, K) in shapes:
    x = torch.randn(1, 1, M, K, generator=g).cuda()
    w = torch.randn(N, K, 1, 1, generator=g) / K ** 0.5
    hi, lo = U.split_hi_lo(w.reshape(N, K)); hi, lo = hi.cuda(), lo.cuda()
    b = torch.randn(N, generator=g).cuda()
    y = torch.empty(1, 1, M, N, device="cuda")
    for pair in (0, 1):
        if pair: os.environ.pop("PF_NO_PAIR", None)
        else: os.environ["PF_NO_PAIR"] = "1"
        for dbg in (DBG if not pair else [0]):
            os.environ["PF_GEMM_DBG"] = str(dbg)
            for rep in range(2):
                _native.check(L.pf_op_conv_gemm(x.data_ptr(), 1, 1, M, K, hi.data_ptr(), lo.data_ptr(), b.data_ptr(), N, 1, 1, 1, 0, 0, 0, None, 0, y.data_ptr(), U.stream_ptr()))
            print(f"M{M} N{N} K{K} pair={pair} dbg={dbg}: 2 launches", flush=True)
os.environ.pop("PF_GEMM_DBG", None)
# pure-write / copy bandwidth of the part, for the store-bound kernels' roofline
for mb in (128, 1024):
    n = mb << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda"); c = torch.empty_like(a)
    for fn, name, bytes_ in ((lambda: a.fill_(1), "fill", n), (lambda: c.copy_(a), "copy", 2 * n)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name} {mb} MiB: {bytes_ * 10 / e0.elapsed_time(e1) / 1e6:.0f} GB/s", flush=True)
