"""Drop-in for ``perspective2d.PerspectiveFields`` (reference: perspective2d/perspectivefields.py:121-272) whose
whole forward runs in libpf_b200.so (hand-written sm_100a CUDA, C ABI in include/pf_b200.h).

Kept from the reference surface: ``PerspectiveFields(version)``, ``.eval()``, ``.cuda()/.to()``, ``.device``,
``.versions()``, ``.inference(img_bgr)``, ``.inference_batch(list)``, ``.forward(batched_inputs)``,
``.state_dict()/.load_state_dict()`` with the reference's key names, attributes ``version``, ``param_on``, ``cfg``,
``input_format``; result dictionaries with the same keys, order, shapes and dtypes.  There is no CPU path: the model
must live on a CUDA device (B200) and libpf_b200.so must be built, otherwise inference raises.
"""
import ctypes

import numpy as np
import torch
from torch import nn

from . import _native
from .checkpoint import checkpoint_schema, default_state, load_zoo_checkpoint
from .variants import PIXEL_MEAN, PIXEL_STD, RESIZE, VARIANTS, make_cfg, model_zoo
from .weights import repack

_NET = RESIZE[0]


class _Engine:
    """One libpf_b200 handle + its device-resident repacked weights and scratch, for one CUDA device."""

    def __init__(self, device, version, ref_state):
        self.L = _native.lib()
        self.device = device
        cfg = VARIANTS[version]
        desc = _native.pf_model_desc()
        desc.gravity_classes, desc.latitude_classes = cfg["gravity_classes"], cfg["latitude_classes"]
        desc.param_net = {None: _native.PF_PARAM_NONE, "ParamNet": _native.PF_PARAM_CENTERED,
                          "ParamNetConvNextRegress": _native.PF_PARAM_UNCENTERED}[cfg["param_net"]]
        desc.param_input_size = cfg["input_size"]
        desc.pixel_mean[:] = PIXEL_MEAN
        desc.pixel_std[:] = PIXEL_STD
        self.handle = ctypes.c_void_p()
        _native.check(self.L.pf_create(device.index, ctypes.byref(desc), ctypes.byref(self.handle)))
        self.tensors = {}
        for name, t in repack(ref_state, cfg).items():
            d = t.to(device)
            self.tensors[name] = d  # keeps the device memory alive for the lifetime of the handle
            dt = _native.PF_BF16 if d.dtype == torch.bfloat16 else _native.PF_F32
            _native.check(self.L.pf_set_weight(self.handle, name.encode(), d.data_ptr(), d.numel(), dt))
        _native.check(self.L.pf_finalize(self.handle))
        self.workspace = None
        self.ws_stream = None      # stream of the last forward that used the workspace
        self.ws_event = None       # ... and its completion
        self.decode_only = False
        self.pinned = None
        self.pinned_event = None
        self.dev_blob = None
        self.staged_slot = None

    def close(self):
        if self.handle:
            self.L.pf_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _workspace(self, n, max_h, cur):
        """Scratch for pf_forward.  The buffer is allocated from torch's caching allocator on the stream of its first use; a
        caller that switches streams between calls is kept safe by stream-ordering the hand-over: the new stream waits for the
        last forward that used the workspace (``ws_event``), and a workspace that is replaced is marked as used by that stream
        (``record_stream``) so that its block is not recycled under a forward still running there."""
        need = _native.check(self.L.pf_workspace_bytes(self.handle, n, max_h))
        if self.ws_event is not None and self.ws_stream is not None and self.ws_stream != cur:
            cur.wait_event(self.ws_event)
        if self.workspace is None or self.workspace.numel() < need:
            if self.workspace is not None and self.ws_stream is not None:
                self.workspace.record_stream(self.ws_stream)
            self.workspace = None
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        self.ws_stream = cur
        return self.workspace

    def stage_images(self, imgs):
        """Host uint8 images -> one pinned blob -> one async H2D copy on a dedicated copy stream (two pinned / device blob pairs,
        so the upload of batch k+1 overlaps the forward of batch k).  Returns (device blob, offsets); the current stream has been
        made to wait for the upload."""
        sizes = [im.size for im in imgs]
        offsets = np.zeros(len(imgs), np.int64)
        np.cumsum(sizes[:-1], out=offsets[1:])
        total = int(sum(sizes))
        if self.pinned is None or self.pinned[0].numel() < total:
            cap = max(total, 1 << 20)
            self.pinned = [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(2)]
            self.dev_blob = [torch.empty(cap, dtype=torch.uint8, device=self.device) for _ in range(2)]
            self.pinned_event = [None, None]    # upload from pinned[s] has completed
            self.blob_free = [None, None]       # the forward that read dev_blob[s] has completed (recorded by forward())
            self.h2d_stream = torch.cuda.Stream(device=self.device)
            self.slot = 0
        s = self.slot = self.slot ^ 1
        if self.pinned_event[s] is not None:
            self.pinned_event[s].synchronize()  # the upload that last used this pinned buffer must have drained before it is rewritten
        host = self.pinned[s].numpy()
        for im, off in zip(imgs, offsets):
            host[off:off + im.size] = im.reshape(-1)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.h2d_stream):
            if self.blob_free[s] is not None:
                self.h2d_stream.wait_event(self.blob_free[s])
            else:
                self.h2d_stream.wait_stream(cur)
            self.dev_blob[s][:total].copy_(self.pinned[s][:total], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.h2d_stream)
        self.pinned_event[s] = ev
        cur.wait_event(ev)
        self.staged_slot = s
        return self.dev_blob[s], offsets

    def forward(self, n, heights, widths, blob=None, offsets=None, chw=None):
        dev = self.device
        h = np.ascontiguousarray(heights, np.int32)
        w = np.ascontiguousarray(widths, np.int32)
        hw = h.astype(np.int64) * w.astype(np.int64)
        g_off = np.zeros(n, np.int64)
        l_off = np.zeros(n, np.int64)
        np.cumsum(2 * hw[:-1], out=g_off[1:])
        np.cumsum(hw[:-1], out=l_off[1:])
        cur = torch.cuda.current_stream(dev)
        gc_, lc_ = (2, 1) if self.decode_only else (self.gravity_classes, self.latitude_classes)
        out = {
            "pred_gravity": torch.empty((n, gc_, _NET, _NET), dtype=torch.float32, device=dev),
            "pred_latitude": torch.empty((n, lc_, _NET, _NET), dtype=torch.float32, device=dev),
            "gravity_original": torch.empty(int(2 * hw.sum()), dtype=torch.float32, device=dev),
            "latitude_original": torch.empty(int(hw.sum()), dtype=torch.float32, device=dev),
            "params": torch.empty((n, 8), dtype=torch.float32, device=dev),
            "g_off": g_off, "l_off": l_off, "h": h, "w": w,
        }
        ws = self._workspace(n, int(h.max()), cur)
        bt = _native.pf_batch()
        bt.n = n
        i64p, i32p = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32)
        if blob is not None:
            offsets = np.ascontiguousarray(offsets, np.int64)
            bt.images_u8 = blob.data_ptr()
            bt.image_offset = offsets.ctypes.data_as(i64p)
        else:
            bt.images_chw = chw.data_ptr()
        bt.height, bt.width = h.ctypes.data_as(i32p), w.ctypes.data_as(i32p)
        bt.pred_gravity, bt.pred_latitude = out["pred_gravity"].data_ptr(), out["pred_latitude"].data_ptr()
        bt.gravity_original, bt.gravity_original_offset = out["gravity_original"].data_ptr(), g_off.ctypes.data_as(i64p)
        bt.latitude_original, bt.latitude_original_offset = out["latitude_original"].data_ptr(), l_off.ctypes.data_as(i64p)
        bt.params = out["params"].data_ptr()
        _native.check(self.L.pf_forward(self.handle, ctypes.byref(bt), ws.data_ptr(), ws.numel(), cur.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(cur)
        self.ws_event = ev
        if blob is not None and self.staged_slot is not None and self.dev_blob is not None and blob is self.dev_blob[self.staged_slot]:
            self.blob_free[self.staged_slot] = ev   # stage_images may overwrite this device blob once the forward has read it
        return out


class ResizeTransform:
    """The ``aug`` attribute of the reference class (perspectivefields.py:16-67, built at :155).  ``apply_image`` keeps the
    reference's contract -- numpy HWC in, numpy HWC out, uint8 through Pillow's antialiased bilinear resampler (bit-exact
    integer restatement, csrc/prepost.cuh: the same arithmetic ``inference`` uses inside its fused pre-process), any other
    dtype through ``F.interpolate(mode="bilinear", align_corners=False)`` -- but the arithmetic runs on the GPU
    (``pf_op_resize_u8`` / ``pf_op_resize_f32``).  Only the bilinear filter exists here (the one the path uses)."""

    def __init__(self, new_h, new_w, interp=None):
        self.new_h, self.new_w = new_h, new_w
        self.interp = 2 if interp is None else interp          # PIL.Image.BILINEAR == 2

    def apply_image(self, img, interp=None):
        img = np.asarray(img)
        assert len(img.shape) <= 4
        method = self.interp if interp is None else interp
        if method != 2:
            raise NotImplementedError("perspectivefields_b200 implements the BILINEAR resize of the inference path only")
        if not torch.cuda.is_available():
            raise RuntimeError("perspectivefields_b200 has no CPU path: ResizeTransform.apply_image needs a CUDA device")
        L = _native.lib()
        dev = torch.device("cuda", torch.cuda.current_device())
        stream = torch.cuda.current_stream(dev).cuda_stream
        if img.dtype == np.uint8:
            if img.ndim != 3 or img.shape[2] != 3:
                raise TypeError("uint8 images must be (H, W, 3); got %s" % (img.shape,))
            src = torch.from_numpy(np.ascontiguousarray(img)).to(dev)
            out = torch.empty((self.new_h, self.new_w, 3), dtype=torch.uint8, device=dev)
            _native.check(L.pf_op_resize_u8(src.data_ptr(), img.shape[0], img.shape[1], self.new_h, self.new_w, out.data_ptr(), stream))
            return out.cpu().numpy()
        if img.ndim not in (2, 3):
            raise TypeError("float images must be (H, W) or (H, W, C); got %s" % (img.shape,))
        c = 1 if img.ndim == 2 else img.shape[2]
        src = torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).to(dev)
        out = torch.empty((self.new_h, self.new_w) + img.shape[2:], dtype=torch.float32, device=dev)
        _native.check(L.pf_op_resize_f32(src.data_ptr(), img.shape[0], img.shape[1], c, self.new_h, self.new_w, out.data_ptr(), stream))
        return out.cpu().numpy().astype(img.dtype, copy=False)


class PerspectiveFields(nn.Module):
    def __init__(self, version="Paramnet-360Cities-edina-centered", logits=True):
        """``logits=False`` (classification variant only; NOT the reference's behaviour): ``pred_gravity`` /
        ``pred_latitude`` hold the decoded fields ([2,320,320] up-vectors, [1,320,320] degrees) instead of the 73 / 180 raw logits,
        which are then never written (engine option "decode_only"); the ``*_original`` entries are unchanged."""
        super().__init__()
        zoo = model_zoo[version]  # KeyError for unknown versions, like the reference (perspectivefields.py:127)
        self.version = version
        self.param_on = zoo["param"]
        self.cfg = make_cfg(version)
        self._variant = VARIANTS[version]
        self.register_buffer("pixel_mean", torch.tensor(PIXEL_MEAN).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(PIXEL_STD).view(-1, 1, 1), False)
        self.vis_period = self.cfg.VIS_PERIOD
        self.freeze = self.cfg.MODEL.FREEZE
        self.debug_on = self.cfg.DEBUG_ON
        self.input_format = self.cfg.INPUT.FORMAT
        self.aug = ResizeTransform(RESIZE[0], RESIZE[1])
        self._schema = dict(checkpoint_schema(version))
        self._ref_state = default_state(version)   # reference-layout weights, host side
        self._engine = None
        self._options = {}
        self._jpeg = None
        if not logits:
            if self._variant["gravity"] != "classification":
                raise ValueError("logits=False only applies to the classification variant (PersNet-360Cities)")
            self._options["decode_only"] = 1
        self.training = False
        self._init_weights()

    # ------------------------------------------------------------------------------------------ module plumbing
    @property
    def device(self):
        return self.pixel_mean.device

    @staticmethod
    def versions():
        for key in model_zoo:
            print(f"{key}")
            print(f"   - {model_zoo[key]['description']}")

    def train(self, mode=True):
        if mode:
            raise RuntimeError("perspectivefields_b200.PerspectiveFields is inference-only: call .eval()")
        return super().train(False)

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        """The reference's key layout (backbone.*, ll_enc.*, persformer_heads.*, param_net.backbone.*), with
        ``nn.Module.state_dict``'s arguments: entries are added to ``destination`` under ``prefix``; ``keep_vars`` returns the
        stored tensors themselves instead of detached copies."""
        if args:   # legacy positional form: (destination, prefix, keep_vars)
            destination = args[0]
            prefix = args[1] if len(args) > 1 else prefix
            keep_vars = args[2] if len(args) > 2 else keep_vars
        if destination is None:
            import collections
            destination = collections.OrderedDict()
        for k, v in self._ref_state.items():
            destination[prefix + k] = v if keep_vars else v.detach().clone()
        return destination

    def load_state_dict(self, state_dict, strict=True, assign=False):
        missing = [k for k in self._schema if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._schema]
        errors = []
        for k, v in state_dict.items():
            if k in self._schema:
                if tuple(v.shape) != tuple(self._schema[k]):
                    errors.append(f"size mismatch for {k}: checkpoint {tuple(v.shape)} vs model {tuple(self._schema[k])}")
        if strict and (missing or unexpected):
            errors.append(f"missing keys {missing[:5]}..., unexpected keys {unexpected[:5]}...")
        if errors:
            raise RuntimeError("Error(s) in loading state_dict for PerspectiveFields:\n\t" + "\n\t".join(errors))
        for k, v in state_dict.items():
            if k in self._schema:
                self._ref_state[k] = v.detach().to("cpu", self._ref_state[k].dtype).clone()
        self._drop_engine()
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def _init_weights(self):
        """perspectivefields.py:178-192."""
        state_dict = load_zoo_checkpoint(model_zoo[self.version]["weights"])
        self.load_state_dict(state_dict, strict=False)  # a no-op on the {"model": ...} wrapper, as in the reference
        if state_dict:
            self.load_state_dict(state_dict["model"], strict=False)

    def _drop_engine(self):
        if self._jpeg is not None and self._engine is not None:
            self._engine.L.pf_jpeg_destroy(self._jpeg)
        self._jpeg = None
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def _get_engine(self):
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("perspectivefields_b200 has no CPU path: move the model to a B200 with .cuda() first")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if self._engine is None or self._engine.device != dev:
            self._drop_engine()
            with torch.cuda.device(dev):
                eng = _Engine(dev, self.version, self._ref_state)
            eng.gravity_classes = self._variant["gravity_classes"]
            eng.latitude_classes = self._variant["latitude_classes"]
            for k, v in self._options.items():
                _native.check(eng.L.pf_set_option(eng.handle, k.encode(), v))
            eng.decode_only = bool(self._options.get("decode_only", 0))
            self._engine = eng
        return self._engine

    # ------------------------------------------------------------------------------------------ inference API
    @torch.no_grad()
    def inference(self, img_bgr):
        return self.inference_batch([img_bgr])[0]

    @torch.no_grad()
    def inference_batch(self, img_bgr_list):
        """perspectivefields.py:207-221.  uint8 (H, W, 3) images take the fused path (one packed upload, Pillow-exact resize +
        normalise in one kernel); a list containing any other dtype takes the reference's float branch for ALL its members
        (``ResizeTransform.apply_image`` -> non-antialiased ``F.interpolate``, perspectivefields.py:47-66, on the GPU) and then
        the ``forward`` entry."""
        imgs = []
        all_u8 = True
        for im in img_bgr_list:
            im = np.asarray(im)
            if im.ndim != 3 or im.shape[2] != 3:
                raise TypeError("inference expects (H, W, 3) BGR images; got %s %s" % (im.dtype, im.shape))
            if self.input_format == "RGB":
                im = im[:, :, ::-1]
            all_u8 = all_u8 and im.dtype == np.uint8
            imgs.append(np.ascontiguousarray(im))
        if not imgs:
            return []
        eng = self._get_engine()
        with torch.cuda.device(eng.device):
            if not all_u8:
                inputs = []
                for im in imgs:
                    r = self.aug.apply_image(im)
                    inputs.append({"image": torch.as_tensor(r.astype("float32").transpose(2, 0, 1)), "height": im.shape[0], "width": im.shape[1]})
                return self.forward(inputs)
            blob, offsets = eng.stage_images(imgs)
            out = eng.forward(len(imgs), [im.shape[0] for im in imgs], [im.shape[1] for im in imgs], blob=blob, offsets=offsets)
        return self._assemble(out)

    # ---- blob-level access for the multi-GPU gather (dist.py): the five output blobs of a batch move as whole buffers ----
    @torch.no_grad()
    def infer_raw(self, img_bgr_list):
        """``inference_batch`` up to (not including) the per-image views: returns the engine's batch outputs
        (``pred_gravity [n,Cg,320,320]``, ``pred_latitude``, flat ``gravity_original`` / ``latitude_original`` blobs, ``params [n,8]``)
        plus the host-side offsets / sizes ``assemble_raw`` needs.  uint8 images only."""
        imgs = []
        for im in img_bgr_list:
            im = np.asarray(im)
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise TypeError("infer_raw expects (H, W, 3) uint8 BGR images; got %s %s" % (im.dtype, im.shape))
            if self.input_format == "RGB":
                im = im[:, :, ::-1]
            imgs.append(np.ascontiguousarray(im))
        eng = self._get_engine()
        with torch.cuda.device(eng.device):
            blob, offsets = eng.stage_images(imgs)
            return eng.forward(len(imgs), [im.shape[0] for im in imgs], [im.shape[1] for im in imgs], blob=blob, offsets=offsets)

    def assemble_raw(self, raw):
        return self._assemble(raw)

    def out_classes(self):
        """Channel counts of ``pred_gravity`` / ``pred_latitude`` as returned (2 / 1 in "decode_only" mode)."""
        if self._options.get("decode_only", 0):
            return (2, 1)
        return (self._variant["gravity_classes"], self._variant["latitude_classes"])

    def decode_batch(self, jpeg_list, max_threads=0):
        """Decode front-end: a list of JPEG byte strings (what ``cv2.imread`` would read from disk,
        demo/demo.py:151) is decoded on the GPU (nvJPEG, BGR interleaved) into ONE device blob of packed HWC uint8 images -- the
        layout the fused pre-process reads; no host-side pixel buffer exists.  Returns (blob, offsets, heights, widths)."""
        if self.input_format != "BGR":
            raise NotImplementedError("the decode front-end writes BGR (the reference's INPUT.FORMAT)")
        eng = self._get_engine()
        L = eng.L
        with torch.cuda.device(eng.device):
            if self._jpeg is None:
                self._jpeg = ctypes.c_void_p()
                _native.check(L.pf_jpeg_create(eng.device.index, max_threads, ctypes.byref(self._jpeg)))
            n = len(jpeg_list)
            bufs = [np.frombuffer(b, dtype=np.uint8) for b in jpeg_list]
            hs, ws = (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)()
            for i, b in enumerate(bufs):
                h1, w1 = ctypes.c_int32(), ctypes.c_int32()
                _native.check(L.pf_jpeg_info(self._jpeg, b.ctypes.data, b.size, ctypes.byref(h1), ctypes.byref(w1)))
                hs[i], ws[i] = h1.value, w1.value
            sizes = [hs[i] * ws[i] * 3 for i in range(n)]
            offsets = np.zeros(n, np.int64)
            np.cumsum(sizes[:-1], out=offsets[1:])
            offs = (ctypes.c_int64 * n)(*offsets.tolist())
            blob = torch.empty(int(sum(sizes)), dtype=torch.uint8, device=eng.device)
            ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
            lens = (ctypes.c_int64 * n)(*[b.size for b in bufs])
            stream = torch.cuda.current_stream(eng.device).cuda_stream
            _native.check(L.pf_jpeg_decode_batch(self._jpeg, n, ptrs, lens, hs, ws, blob.data_ptr(), offs, stream))
        return blob, offsets, list(hs), list(ws)

    @torch.no_grad()
    def inference_batch_encoded(self, jpeg_list, max_threads=0):
        """``inference_batch`` on JPEG byte strings: ``decode_batch`` + the forward on the decoded blob.  Returns the same
        ``list[dict]`` as ``inference_batch`` (up to the decoder: nvJPEG's IDCT / chroma up-sampling round differently from
        libjpeg's, so decoded pixels can differ by a few grey levels from ``cv2.imread``)."""
        if not jpeg_list:
            return []
        blob, offsets, hs, ws = self.decode_batch(jpeg_list, max_threads)
        eng = self._get_engine()
        with torch.cuda.device(eng.device):
            out = eng.forward(len(jpeg_list), hs, ws, blob=blob, offsets=offsets)
        return self._assemble(out)

    @torch.no_grad()
    def forward(self, batched_inputs):
        """perspectivefields.py:223-272: ``[{"image": float32 [3,320,320] (resized, un-normalised), "height", "width"}]``."""
        if any(k in batched_inputs[0] for k in ("gt_gravity", "gt_latitude")) and self.training:
            raise RuntimeError("training is not supported")
        eng = self._get_engine()
        with torch.cuda.device(eng.device):
            chw = torch.stack([x["image"].to(eng.device, torch.float32) for x in batched_inputs]).contiguous()
            if tuple(chw.shape[1:]) != (3, _NET, _NET):
                raise ValueError("forward expects images already resized to [3, 320, 320]")
            out = eng.forward(len(batched_inputs), [int(x["height"]) for x in batched_inputs],
                              [int(x["width"]) for x in batched_inputs], chw=chw)
        return self._assemble(out)

    def _assemble(self, out):
        """Result dictionaries: keys and order of persformer_heads.py:83-101 + param_network.py:54-67 / 205-220 +
        perspectivefields.py:261-271."""
        v = self._variant
        n = out["pred_gravity"].shape[0]
        # one unbind / split call per output tensor (not ~12 tensor operations per image: at 256 images per call the per-image
        # Python work was several milliseconds)
        hs, ws = [int(x) for x in out["h"]], [int(x) for x in out["w"]]
        pg, pl = out["pred_gravity"].unbind(0), out["pred_latitude"].unbind(0)
        if len(set(zip(hs, ws))) == 1:
            go_ = out["gravity_original"].view(n, 2, hs[0], ws[0]).unbind(0)
            lo_ = out["latitude_original"].view(n, hs[0], ws[0]).unbind(0)
        else:
            go_ = [t.view(2, h, w) for t, h, w in zip(out["gravity_original"].split([2 * h * w for h, w in zip(hs, ws)]), hs, ws)]
            lo_ = [t.view(h, w) for t, h, w in zip(out["latitude_original"].split([h * w for h, w in zip(hs, ws)]), hs, ws)]
        cols = [c.unbind(0) for c in out["params"].t().unbind(0)] if v["param_net"] else None     # cols[j][i] = params[i, j] (0-dim views)
        zeros = torch.zeros_like(out["params"][:, 0]).unbind(0) if v["param_net"] == "ParamNet" else None
        res = []
        for i in range(n):
            d = {"pred_gravity": pg[i], "pred_gravity_original": go_[i], "pred_latitude": pl[i], "pred_latitude_original": lo_[i],
                 "pred_latitude_original_mode": "deg"}
            if v["param_net"] == "ParamNet":
                d.update({"pred_roll": cols[0][i], "pred_pitch": cols[1][i], "pred_vfov": cols[2][i], "pred_rel_focal": cols[5][i],
                          "pred_general_vfov": cols[2][i], "pred_rel_cx": zeros[i], "pred_rel_cy": zeros[i]})
            elif v["param_net"] == "ParamNetConvNextRegress":
                d.update({"pred_roll": cols[0][i], "pred_pitch": cols[1][i], "pred_general_vfov": cols[2][i], "pred_rel_cx": cols[3][i],
                          "pred_rel_cy": cols[4][i], "pred_rel_focal": cols[5][i]})
            res.append(d)
        return res

    def set_option(self, name, value):
        """Engine options (see pf_set_option in include/pf_b200.h), e.g. ``set_option("tcgen05", 1)``."""
        eng = self._get_engine()
        _native.check(eng.L.pf_set_option(eng.handle, name.encode(), int(value)))
        self._options[name] = int(value)
        eng.decode_only = bool(self._options.get("decode_only", 0))

    # ------------------------------------------------------------------------------------------ test hooks
    def debug_taps(self, enable=True):
        eng = self._get_engine()
        _native.check(eng.L.pf_debug_enable(eng.handle, 1 if enable else 0))
        eng.workspace = None

    def read_taps(self):
        eng = self._get_engine()
        L, out = eng.L, {}
        stream = torch.cuda.current_stream(eng.device).cuda_stream
        for i in range(L.pf_debug_count(eng.handle)):
            name = L.pf_debug_name(eng.handle, i)
            numel = L.pf_debug_numel(eng.handle, name)
            t = torch.empty(numel, dtype=torch.float32, device=eng.device)
            _native.check(L.pf_debug_copy(eng.handle, name, t.data_ptr(), numel, stream))
            out[name.decode()] = t
        torch.cuda.synchronize(eng.device)
        return out
