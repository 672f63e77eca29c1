"""GPU, two devices, NCCL: ``dist.inference_batch_sharded`` on real engines -- every rank passes the whole list, each runs its shard
in micro-batches, the five output blobs of every micro-batch travel to rank 0 with ``pf_gather`` (grouped ncclSend / ncclRecv
issued from libpf_b200.so on a side stream).  The gathered ``list[dict]`` must equal a single-GPU ``inference_batch`` of the same
list BIT FOR BIT (batch composition never changes an image's result; the transport moves bytes)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, version, q):
    import torch.distributed as dist

    import pf_test_util as U
    from oracle import weights_gen as wg
    from perspectivefields_b200 import dist as pfdist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        model, _ = U.make_model(version, device=dev)
        imgs = (wg.smooth_images(2, 120, 160, 5) + wg.synth_images(1, 96, 200, 6) + wg.smooth_images(2, 240, 320, 7) + wg.synth_images(2, 64, 64, 8))
        tr = pfdist.PfCommTransport(dev)
        ok = True
        for mb in (2, 32):
            res = pfdist.inference_batch_sharded(model, imgs, gather_to=0, micro_batch=mb, transport=tr)
            torch.cuda.synchronize(dev)
            if rank == 0:
                ref = model.inference_batch(imgs)
                ok = ok and len(res) == len(ref) == len(imgs)
                for a, b in zip(res, ref):
                    ok = ok and list(a.keys()) == list(b.keys())
                    for k, v in b.items():
                        if isinstance(v, str):
                            ok = ok and a[k] == v
                        else:
                            ok = ok and a[k].device == v.device and torch.equal(a[k], v)
            else:
                lo, hi = pfdist.shard_bounds(len(imgs), world)[rank]
                ok = ok and len(res) == hi - lo
        moved = tr.bytes_moved
        tr.close()
        q.put((rank, bool(ok), int(moved)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("version", ["PersNet_Paramnet-GSV-uncentered"])
def test_sharded_inference_two_gpus_nccl_gather_is_bit_identical(version):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, version, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    assert [g[:2] for g in got] == [(0, True), (1, True)], got
    assert got[0][2] > 0 and got[1][2] > 0          # bytes really moved through pf_gather on both sides
