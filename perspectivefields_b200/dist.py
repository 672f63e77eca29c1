"""Multi-GPU ``inference_batch``: one process per GPU (``torch.distributed``, NCCL over NVLink on a B200 box; gloo in the CPU
tests), images sharded in contiguous chunks, no data-path collective -- the path has no cross-image dependency (SURVEY.md
section 8e).  Optionally the per-image results are gathered on one rank with point-to-point sends so that the caller gets
the same ``list[dict]`` a single-GPU call would return (bit-identical per image: batch composition never changes a result).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world):
    """Contiguous shards of ceil(n / world) images: [(lo, hi)] per rank (empty shards at the tail are legal)."""
    per = -(-n // world) if n else 0
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def _result_spec(variant, h, w):
    """(key, shape) of every tensor in one image's result dict, in order."""
    g, l = variant["gravity_classes"], variant["latitude_classes"]
    spec = [("pred_gravity", (g, 320, 320)), ("pred_gravity_original", (2, h, w)), ("pred_latitude", (l, 320, 320)),
            ("pred_latitude_original", (h, w))]
    if variant["param_net"] == "ParamNet":
        spec += [(k, ()) for k in ("pred_roll", "pred_pitch", "pred_vfov", "pred_rel_focal", "pred_general_vfov", "pred_rel_cx", "pred_rel_cy")]
    elif variant["param_net"] == "ParamNetConvNextRegress":
        spec += [(k, ()) for k in ("pred_roll", "pred_pitch", "pred_general_vfov", "pred_rel_cx", "pred_rel_cy", "pred_rel_focal")]
    return spec


def inference_batch_sharded(model, img_bgr_list, gather_to=0, group=None):
    """Every rank passes the SAME list; rank r runs ``model.inference_batch`` on its shard.  With ``gather_to`` = a rank, that
    rank returns the full ``list[dict]`` in input order (tensors on its device) and the others return their own shard's
    results; with ``gather_to=None`` nothing is exchanged."""
    if not dist.is_available() or not dist.is_initialized():
        return model.inference_batch(img_bgr_list)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    bounds = shard_bounds(len(img_bgr_list), world)
    lo, hi = bounds[rank]
    mine = model.inference_batch(img_bgr_list[lo:hi]) if hi > lo else []
    if gather_to is None or world == 1:
        return mine
    variant = model._variant
    device = mine[0]["pred_gravity"].device if mine else model.device
    ops, keep = [], []
    if rank == gather_to:
        full = [None] * len(img_bgr_list)
        for i, d in zip(range(lo, hi), mine):
            full[i] = d
        for r, (a, b) in enumerate(bounds):
            if r == rank:
                continue
            for i in range(a, b):
                h, w = img_bgr_list[i].shape[:2]
                d = {}
                for k, shape in _result_spec(variant, h, w):
                    t = torch.empty(shape, dtype=torch.float32, device=device)
                    d[k] = t
                    ops.append(dist.P2POp(dist.irecv, t, r, group))
                # keep the reference key order, including the string entry
                ordered = {}
                for k in d:
                    ordered[k] = d[k]
                    if k == "pred_latitude_original":
                        ordered["pred_latitude_original_mode"] = "deg"
                full[i] = ordered
    else:
        for d in mine:
            for k, v in d.items():
                if isinstance(v, str):
                    continue
                t = v.contiguous()
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, gather_to, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return full if rank == gather_to else mine
