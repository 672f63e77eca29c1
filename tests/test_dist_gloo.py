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


def test_micro_batches_and_blob_sizes():
    assert pfdist.micro_batches(3, 10, 4) == [(3, 7), (7, 10)] and pfdist.micro_batches(5, 5, 4) == []
    n = pfdist.blob_numels((2, 1), [(480, 640), (10, 20)])
    assert n["pred_gravity"] == 2 * 2 * 320 * 320 and n["gravity_original"] == 2 * (480 * 640 + 200) and n["params"] == 16
    raw = pfdist.empty_raw((73, 180), [(4, 6), (3, 5)], "cpu")
    assert raw["pred_latitude"].shape == (2, 180, 320, 320) and list(raw["g_off"]) == [0, 48] and list(raw["l_off"]) == [0, 24]


def test_shard_bounds():
    assert pfdist.shard_bounds(256, 8) == [(i * 32, (i + 1) * 32) for i in range(8)]
    assert pfdist.shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert pfdist.shard_bounds(1, 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]
    assert pfdist.shard_bounds(0, 2) == [(0, 0), (0, 0)]


class _FakeModel:
    """Same blob-level contract as PerspectiveFields (infer_raw / assemble_raw / out_classes / inference_batch) on CPU tensors;
    values are a function of the image content only."""

    def __init__(self, version):
        self._variant = VARIANTS[version]
        self.device = torch.device("cpu")
        self.calls = 0

    def out_classes(self):
        return (self._variant["gravity_classes"], self._variant["latitude_classes"])

    def infer_raw(self, imgs):
        self.calls += len(imgs)
        raw = pfdist.empty_raw(self.out_classes(), [im.shape[:2] for im in imgs], self.device)
        for i, im in enumerate(imgs):
            tag = float(im[0, 0, 0])
            h, w = im.shape[:2]
            raw["pred_gravity"][i] = tag + 1
            raw["pred_latitude"][i] = tag + 2
            raw["gravity_original"][raw["g_off"][i]:raw["g_off"][i] + 2 * h * w] = tag + 3
            raw["latitude_original"][raw["l_off"][i]:raw["l_off"][i] + h * w] = tag + 4
            raw["params"][i] = tag + torch.arange(8, dtype=torch.float32)
        return raw

    def assemble_raw(self, raw):
        out = []
        for i in range(raw["pred_gravity"].shape[0]):
            h, w = int(raw["h"][i]), int(raw["w"][i])
            go, lo = int(raw["g_off"][i]), int(raw["l_off"][i])
            d = {"pred_gravity": raw["pred_gravity"][i], "pred_gravity_original": raw["gravity_original"][go:go + 2 * h * w].view(2, h, w),
                 "pred_latitude": raw["pred_latitude"][i], "pred_latitude_original": raw["latitude_original"][lo:lo + h * w].view(h, w),
                 "pred_latitude_original_mode": "deg"}
            if self._variant["param_net"] is not None:
                d["pred_roll"] = raw["params"][i, 0]
            out.append(d)
        return out

    def inference_batch(self, imgs):
        return self.assemble_raw(self.infer_raw(imgs))


def _worker(rank, world, port, version, sizes, mb, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        imgs = [np.full((h, w, 3), i + 1, np.uint8) for i, (h, w) in enumerate(sizes)]
        model = _FakeModel(version)
        res = pfdist.inference_batch_sharded(model, imgs, gather_to=0, micro_batch=mb)
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


@pytest.mark.parametrize("version,sizes,mb", [("Paramnet-360Cities-edina-centered", [(48, 64), (30, 50), (64, 48), (20, 20), (33, 47)], 32),
                                              ("Paramnet-360Cities-edina-centered", [(48, 64), (30, 50), (64, 48), (20, 20), (33, 47)], 2),   # ragged rounds
                                              ("PersNet-360Cities", [(24, 32)], 1)])
def test_sharded_inference_gloo_world2(version, sizes, mb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, version, sizes, mb, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert got == [(0, True), (1, True)]
