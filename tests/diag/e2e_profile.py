"""Host-side time breakdown of one end-to-end step (bench.py's e2e leg): where does the Python thread spend its time?
usage: python tests/diag/e2e_profile.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import pf_test_util as U
from oracle import weights_gen as wg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m, sd = U.make_model("Paramnet-360Cities-edina-centered")
imgs = wg.synth_images(B, 480, 640, 0)
eng = m._get_engine()
for _ in range(3):
    r = m.inference_batch(imgs)
torch.cuda.synchronize()
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1000
K = 5
keys = [k for k, v in r[0].items() if not isinstance(v, str)]
host = {k: torch.empty((B,) + tuple(r[0][k].shape), dtype=torch.float32).pin_memory() for k in keys}
for sync in (True, False):
    T.clear()
    t_all = time.perf_counter()
    for _ in range(K):
        t0 = time.perf_counter(); blob, offs = eng.stage_images(imgs); tick("stage_images", t0)
        if sync: torch.cuda.synchronize()
        t0 = time.perf_counter(); out = eng.forward(B, [480] * B, [640] * B, blob=blob, offsets=offs); tick("forward_enqueue", t0)
        if sync: t0 = time.perf_counter(); torch.cuda.synchronize(); tick("gpu_wait", t0)
        t0 = time.perf_counter(); res = m._assemble(out); tick("assemble", t0)
        t0 = time.perf_counter()
        for i, d in enumerate(res):
            for k in keys:
                host[k][i].copy_(d[k], non_blocking=True)
        tick("d2h_352_copies_enqueue", t0)
        t0 = time.perf_counter(); torch.cuda.synchronize(); tick("d2h_wait", t0)
    total = (time.perf_counter() - t_all) * 1000 / K
    print("sync-between-phases" if sync else "no-sync", "total ms/step %.2f" % total, {k: round(v / K, 2) for k, v in T.items()}, flush=True)
# empty-queue enqueue cost of pf_forward alone
torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); out = eng.forward(B, [480] * B, [640] * B, blob=blob, offsets=offs); ts.append((time.perf_counter() - t0) * 1000)
    torch.cuda.synchronize()
print("pf_forward enqueue on an empty queue (ms):", [round(t, 2) for t in ts])

# ---- stall hunt: 60 pipelined steps, per-component host times; report outliers
import gc
copy_stream = torch.cuda.Stream()
host2 = {k: torch.empty_like(v).pin_memory() for k, v in host.items()}
pending, rows = [], []
orig_empty = torch.empty
for it in range(60):
    c = {}
    t0 = time.perf_counter(); blob, offs = eng.stage_images(imgs); c["stage"] = (time.perf_counter() - t0) * 1000
    t_alloc = [0.0]
    def timed_empty(*a, **k):
        t = time.perf_counter(); r_ = orig_empty(*a, **k); t_alloc[0] += (time.perf_counter() - t) * 1000; return r_
    torch.empty = timed_empty
    t0 = time.perf_counter(); out = eng.forward(B, [480] * B, [640] * B, blob=blob, offsets=offs); c["forward"] = (time.perf_counter() - t0) * 1000
    torch.empty = orig_empty
    c["alloc_in_forward"] = t_alloc[0]
    t0 = time.perf_counter(); res = m._assemble(out); c["assemble"] = (time.perf_counter() - t0) * 1000
    done = torch.cuda.Event(); done.record()
    bufs = host if it % 2 == 0 else host2
    t0 = time.perf_counter()
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(done)
        for i, d in enumerate(res):
            for k in keys:
                bufs[k][i].copy_(d[k], non_blocking=True)
                d[k].record_stream(copy_stream)
    c["d2h_enqueue"] = (time.perf_counter() - t0) * 1000
    t0 = time.perf_counter()
    if len(pending) >= 2:
        pending.pop(0).synchronize()
    c["wait_prev"] = (time.perf_counter() - t0) * 1000
    ev = torch.cuda.Event(); ev.record(copy_stream); pending.append(ev)
    rows.append(c)
torch.cuda.synchronize()
for name in rows[0]:
    v = sorted(r_[name] for r_ in rows[5:])
    print("%-18s median %.2f  p90 %.2f  max %.2f" % (name, v[len(v) // 2], v[int(len(v) * 0.9)], v[-1]))
print("outlier steps (>8 ms host in forward):", [(i, round(r_["forward"], 1), round(r_["alloc_in_forward"], 1)) for i, r_ in enumerate(rows) if r_["forward"] > 8])
print("torch allocator:", {k: v for k, v in torch.cuda.memory_stats().items() if k in ("num_alloc_retries", "num_device_alloc", "num_device_free", "reserved_bytes.all.peak")})
