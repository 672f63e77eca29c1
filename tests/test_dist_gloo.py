"""CPU, world_size 2, gloo: the N > 1 host logic -- contiguous sharding, per-rank inference on the shard, point-to-point gather
of variable-size results to rank 0 in input order.  The CUDA engine is replaced by a stand-in model that returns CPU tensors
which encode the image they belong to (the real engine's batch-invariance is covered by tests/test_gpu_forward.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from perspectivefields_b200 import dist as pfdist
from perspectivefields_b200.variants import VARIANTS


def test_shard_bounds():
    assert pfdist.shard_bounds(256, 8) == [(i * 32, (i + 1) * 32) for i in range(8)]
    assert pfdist.shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert pfdist.shard_bounds(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert pfdist.shard_bounds(0, 2) == [(0, 0), (0, 0)]


class _FakeModel:
    """Same result-dict contract as PerspectiveFields; values are a function of the image content only."""

    def __init__(self, version):
        self._variant = VARIANTS[version]
        self.device = torch.device("cpu")
        self.calls = 0

    def inference_batch(self, imgs):
        self.calls += len(imgs)
        out = []
        for im in imgs:
            tag = float(im[0, 0, 0])
            h, w = im.shape[:2]
            d = {}
            for k, shape in pfdist._result_spec(self._variant, h, w):
                d[k] = torch.full(shape, tag + len(k), dtype=torch.float32)
                if k == "pred_latitude_original":
                    d["pred_latitude_original_mode"] = "deg"
            out.append(d)
        return out


def _worker(rank, world, port, version, sizes, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        imgs = [np.full((h, w, 3), i + 1, np.uint8) for i, (h, w) in enumerate(sizes)]
        model = _FakeModel(version)
        res = pfdist.inference_batch_sharded(model, imgs, gather_to=0)
        lo, hi = pfdist.shard_bounds(len(imgs), world)[rank]
        ok = model.calls == hi - lo
        if rank == 0:
            ref = _FakeModel(version).inference_batch(imgs)
            ok = ok and len(res) == len(ref)
            for a, b in zip(res, ref):
                ok = ok and list(a.keys()) == list(b.keys())
                for k, v in b.items():
                    ok = ok and (a[k] == v if isinstance(v, str) else torch.equal(a[k], v))
        else:
            ok = ok and len(res) == hi - lo
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("version,sizes", [("Paramnet-360Cities-edina-centered", [(48, 64), (30, 50), (64, 48), (20, 20), (33, 47)]),
                                           ("PersNet-360Cities", [(24, 32)])])
def test_sharded_inference_gloo_world2(version, sizes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, version, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, True), (1, True)]
