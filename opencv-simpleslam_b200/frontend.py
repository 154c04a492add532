"""Drop-in objects for `lightglue.ALIKED`, `lightglue.LightGlue` and `lightglue.utils.rbd`
as used by `/root/reference/slam/core/features_utils.py:8-9,25-26,94-95,157-162`.
Host code is Python/PyTorch (device memory, streams); all arithmetic runs in libb200slam.so
(hand-written sm_100a kernels) through the C ABI in include/b200slam.h."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, weights as _weights
from ._lib import lib, check


def rbd(data: dict) -> dict:
    """lightglue.utils.rbd: strip the batch dimension (features_utils.py:95,162,240-241)."""
    return {k: v[0] if isinstance(v, (torch.Tensor, np.ndarray, list)) else v for k, v in data.items()}


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("b200slam needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError(f"b200slam runs on CUDA only, got device {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class _Module:
    """The little bit of nn.Module surface the reference touches: eval(), to(), parameters()."""
    _handle = None

    def eval(self):
        return self

    def train(self, mode=False):
        return self

    def to(self, device=None, *a, **k):
        dev = None
        if isinstance(device, (str, torch.device)):
            dev = torch.device(device)
        if dev is not None and dev.type == "cpu":
            raise RuntimeError("b200slam runs on CUDA only (no CPU fallback)")
        if dev is not None and dev.type == "cuda":
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            if idx != self.device.index:
                self._create(torch.device("cuda", idx))
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device) if isinstance(device, int) else (device or "cuda"))

    def parameters(self):
        # features_utils.py:131 reads next(matcher.parameters()).device
        yield self._param

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass


class ALIKED(_Module):
    """lightglue.ALIKED replacement (features_utils.py:25: `ALIKED(max_num_keypoints=...)`)."""

    def __init__(self, model_name="aliked-n16", max_num_keypoints=-1, detection_threshold=0.2, nms_radius=2,
                 resize=1024, weights=None, device=None, seed=0, **_ignored):
        if model_name not in _weights.ALIKED_CFGS:
            raise ValueError(f"unsupported ALIKED model {model_name}")
        self.model_name, self.max_num_keypoints = model_name, int(max_num_keypoints)
        self.detection_threshold, self.nms_radius, self.resize = float(detection_threshold), int(nms_radius), resize
        if weights is None:
            weights, self.weight_source = _weights.load_aliked_state(model_name, seed)
        else:
            self.weight_source = "provided"
        self._blob = _weights.pack_state(weights)
        self.n_limit = self.max_num_keypoints if self.max_num_keypoints > 0 else 20000
        self._create(_require_cuda(device))

    def _create(self, dev):
        self._destroy()
        cfg = _lib.AlikedCfg()
        lib.b2s_aliked_default_cfg(C.byref(cfg))
        cfg.model = 1 if self.model_name == "aliked-n32" else 0
        cfg.max_kp = self.max_num_keypoints
        cfg.det_thresh = self.detection_threshold
        cfg.nms_radius = self.nms_radius
        cfg.resize_long = int(self.resize) if self.resize else 0
        h = C.c_void_p()
        check(lib.b2s_aliked_create(C.byref(cfg), self._blob, len(self._blob), dev.index, C.byref(h)), "b2s_aliked_create")
        self._handle, self.device = h, dev
        self._undistorter = None
        self._param = torch.empty(1, device=dev)
        with torch.cuda.device(dev):
            self._kp = torch.empty((self.n_limit, 2), dtype=torch.float32, device=dev)
            self._desc = torch.empty((self.n_limit, 128), dtype=torch.float32, device=dev)
            self._sc = torch.empty((self.n_limit,), dtype=torch.float32, device=dev)
            self._n = torch.zeros((1,), dtype=torch.int32, device=dev)
            self._n_host = torch.zeros((1,), dtype=torch.int32).pin_memory()

    def _destroy(self):
        if getattr(self, "_handle", None):
            lib.b2s_aliked_destroy(self._handle)
            self._handle = None

    def set_undistort(self, undistorter=None):
        """Attach a `geometry.FrameUndistorter` (or None): `extract_host*` / `feature_extractor` then take the RAW frame
        and run the reference's `cv2.remap(img, mapx, mapy, INTER_LINEAR)` (main_revamped.py:323-324) on the device in
        front of the extractor - one upload, no host round trip.  Keypoints are in undistorted-image pixels."""
        if undistorter is not None and undistorter.device_index != self.device.index:
            raise ValueError("the undistorter lives on another device")
        check(lib.b2s_aliked_set_undistort(self._handle, undistorter._handle if undistorter is not None else None),
              "b2s_aliked_set_undistort")
        self._undistorter = undistorter            # keep the borrowed handle alive
        return self

    # -- device-resident entry: enqueue on the current stream, outputs stay on the GPU -------
    def extract_device(self, img: torch.Tensor, fmt: int, H: int, W: int, row_stride: int = 0):
        """img: CUDA tensor (u8 HWC BGR or f32 CHW RGB). Returns views (kpts, desc, scores, n_dev)
        into handle-owned buffers, valid until the next call."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(lib.b2s_aliked_extract(self._handle, img.data_ptr(), fmt, H, W, row_stride, st, self._kp.data_ptr(),
                                     self._desc.data_ptr(), self._sc.data_ptr(), self._n.data_ptr()), "b2s_aliked_extract")
        return self._kp, self._desc, self._sc, self._n

    def extract_batch_device(self, imgs, fmt: int, H: int, W: int, row_stride: int = 0, lanes: int = 0, out=None):
        """B same-shape CUDA images -> slabs (kpts [B,max_kp,2], desc [B,max_kp,128], scores [B,max_kp], n int32 [B]) on the
        device (b2s_aliked_extract_batch): the frames run concurrently on `lanes` internal extractor lanes forked from /
        joined into the current stream; no host synchronisation.  `out` = (kpts, desc, scores, n) slabs to fill."""
        B = len(imgs)
        dev = self.device
        if out is None:
            out = (torch.empty((B, self.n_limit, 2), dtype=torch.float32, device=dev), torch.empty((B, self.n_limit, 128), dtype=torch.float32, device=dev),
                   torch.empty((B, self.n_limit), dtype=torch.float32, device=dev), torch.zeros((B,), dtype=torch.int32, device=dev))
        ptrs = (C.c_void_p * max(B, 1))(*[t.data_ptr() for t in imgs])
        st = torch.cuda.current_stream(dev).cuda_stream
        assert out[0].is_contiguous() and out[1].is_contiguous() and out[3].is_contiguous()
        check(lib.b2s_aliked_extract_batch(self._handle, ptrs, B, fmt, H, W, row_stride, int(lanes), st, out[0].data_ptr(), out[1].data_ptr(),
                                           out[2].data_ptr() if out[2] is not None else None, out[3].data_ptr()), "b2s_aliked_extract_batch")
        return out

    def _finish(self, H, W):
        self._n_host.copy_(self._n, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        n = int(self._n_host[0])
        if n == _lib.ALIKED_RANGE:
            raise _lib.B2SError(_lib.ALIKED_RANGE_MSG)
        return {"keypoints": self._kp[:n].clone()[None], "descriptors": self._desc[:n].clone()[None],
                "keypoint_scores": self._sc[:n].clone()[None],
                "image_size": torch.tensor([[float(W), float(H)]], device=self.device)}

    @torch.no_grad()
    def extract(self, img: torch.Tensor, **conf) -> dict:
        """upstream Extractor.extract: img f32 [1,3,H,W] or [3,H,W] RGB in [0,1] (features_utils.py:94)."""
        if img.dim() == 3:
            img = img[None]
        assert img.dim() == 4 and img.shape[0] == 1, "batch size must be 1 (as upstream)"
        if img.shape[1] == 1:
            img = img.expand(-1, 3, -1, -1)
        H, W = int(img.shape[-2]), int(img.shape[-1])
        with torch.cuda.device(self.device):
            t = img[0].to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
            self.extract_device(t, _lib.IMG_RGB_F32_CHW, H, W)
            return self._finish(H, W)

    @torch.no_grad()
    def extract_bgr(self, image: np.ndarray) -> dict:
        """u8 BGR HxWx3 image straight from cv2 (fuses `_bgr_to_tensor`, features_utils.py:219-222)."""
        if image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
            raise ValueError("expected an HxWx3 uint8 BGR image")
        H, W = image.shape[:2]
        with torch.cuda.device(self.device):
            t = torch.from_numpy(np.ascontiguousarray(image)).to(self.device, non_blocking=True)
            self.extract_device(t, _lib.IMG_BGR_U8_HWC, H, W, 3 * W)
            return self._finish(H, W)

    def extract_host(self, image: np.ndarray, desc_renorm_eps: float = 0.0):
        """One C-ABI call with host buffers (b2s_aliked_extract_host_ex): returns numpy (kpts, desc, scores).
        desc_renorm_eps > 0 fuses the reference caller's `des /= (||des|| + eps)` (features_utils.py:100)."""
        H, W = image.shape[:2]
        if image.dtype == np.uint8:
            image = np.ascontiguousarray(image); fmt = _lib.IMG_BGR_U8_HWC; stride = 3 * W
        else:
            image = np.ascontiguousarray(image, dtype=np.float32); fmt = _lib.IMG_RGB_F32_CHW; stride = 0
            H, W = image.shape[-2:]
        kp = np.empty((self.n_limit, 2), np.float32)
        de = np.empty((self.n_limit, 128), np.float32)
        sc = np.empty((self.n_limit,), np.float32)
        n = C.c_int32(0)
        check(lib.b2s_aliked_extract_host_ex(self._handle, image.ctypes.data, fmt, H, W, stride, kp.ctypes.data,
                                             de.ctypes.data, sc.ctypes.data, C.addressof(n), float(desc_renorm_eps)),
              "b2s_aliked_extract_host")
        return kp[:n.value], de[:n.value], sc[:n.value]

    def extract_host_split(self, image: np.ndarray, on_keypoints, desc_renorm_eps: float = 0.0):
        """extract_host in three C-ABI calls (begin / keypoints / finish): `on_keypoints(kpts)` runs on the host while
        the descriptor head is still busy on the GPU.  Returns (on_keypoints' result, desc, scores)."""
        H, W = image.shape[:2]
        if image.dtype == np.uint8:
            image = np.ascontiguousarray(image); fmt = _lib.IMG_BGR_U8_HWC; stride = 3 * W
        else:
            image = np.ascontiguousarray(image, dtype=np.float32); fmt = _lib.IMG_RGB_F32_CHW; stride = 0
            H, W = image.shape[-2:]
        check(lib.b2s_aliked_extract_host_begin(self._handle, image.ctypes.data, fmt, H, W, stride, float(desc_renorm_eps)),
              "b2s_aliked_extract_host_begin")
        kp = np.empty((self.n_limit, 2), np.float32)
        n = C.c_int32(0)
        try:
            check(lib.b2s_aliked_extract_host_keypoints(self._handle, kp.ctypes.data, C.addressof(n)), "b2s_aliked_extract_host_keypoints")
            res = on_keypoints(kp[:n.value])
        finally:
            de = np.empty((max(n.value, 1), 128), np.float32)
            sc = np.empty((max(n.value, 1),), np.float32)
            check(lib.b2s_aliked_extract_host_finish(self._handle, de.ctypes.data, sc.ctypes.data), "b2s_aliked_extract_host_finish")
        return res, de[:n.value], sc[:n.value]

    def last_features_device(self, n: int):
        """CUDA copies (kpts [n,2], desc [n,128]) of what the last extract_host* call returned (b2s_aliked_copy_last_features)."""
        with torch.cuda.device(self.device):
            kp = torch.empty((n, 2), dtype=torch.float32, device=self.device)
            de = torch.empty((n, 128), dtype=torch.float32, device=self.device)
            check(lib.b2s_aliked_copy_last_features(self._handle, kp.data_ptr(), de.data_ptr(), int(n),
                                                    torch.cuda.current_stream(self.device).cuda_stream), "b2s_aliked_copy_last_features")
        return kp, de

    def debug(self, name: str) -> np.ndarray:
        n = C.c_size_t(0)
        check(lib.b2s_aliked_debug_get(self._handle, name.encode(), None, 0, C.byref(n)), "debug_get")
        out = np.empty((n.value,), np.float32)
        check(lib.b2s_aliked_debug_get(self._handle, name.encode(), out.ctypes.data, n.value, C.byref(n)), "debug_get")
        return out

    @property
    def launches(self) -> int:
        return int(lib.b2s_aliked_launch_count(self._handle))


class LightGlue(_Module):
    """lightglue.LightGlue replacement (features_utils.py:26: `LightGlue(features='aliked')`)."""

    def __init__(self, features="aliked", depth_confidence=0.95, width_confidence=0.99, filter_threshold=0.1,
                 weights=None, device=None, seed=0, precision="fp32", pruning_threshold=-1, max_kp=2048, **_ignored):
        if features != "aliked":
            raise ValueError("only features='aliked' is supported")
        self.depth_confidence, self.width_confidence = float(depth_confidence), float(width_confidence)
        self.filter_threshold, self.pruning_threshold = float(filter_threshold), int(pruning_threshold)
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.precision, self.max_kp = precision, int(max_kp)
        if weights is None:
            weights, self.weight_source = _weights.load_lightglue_state(seed)
        else:
            self.weight_source = "provided"
        self.n_layers = 1 + max(int(k.split(".")[1]) for k in weights if k.startswith("transformers."))
        self._blob = _weights.pack_state(weights)
        self._create(_require_cuda(device))

    def _create(self, dev):
        self._destroy()
        cfg = _lib.LgCfg()
        lib.b2s_lg_default_cfg(C.byref(cfg))
        cfg.n_layers = self.n_layers
        cfg.depth_conf, cfg.width_conf, cfg.filter_thresh = self.depth_confidence, self.width_confidence, self.filter_threshold
        cfg.pruning_min_kpts = self.pruning_threshold
        cfg.precision = _lib.PRECISIONS[self.precision]
        cfg.max_kp = self.max_kp
        h = C.c_void_p()
        check(lib.b2s_lightglue_create(C.byref(cfg), self._blob, len(self._blob), dev.index, C.byref(h)), "b2s_lightglue_create")
        self._handle, self.device = h, dev
        self._param = torch.empty(1, device=dev)

    def _destroy(self):
        self._fb = None
        if getattr(self, "_handle", None):
            lib.b2s_lg_destroy(self._handle)
            self._handle = None

    # -- precision="fp32" runs on fp16x2 operand planes (include/b200slam.h); a launch sequence in which a value left the
    #    fp16 range reports n = LG_RANGE for its pairs and is re-run on the bf16x3 engine (any range, twice the MMA work)
    def fallback(self):
        """The precision='fp32x3' matcher with the same weights and settings (created on first use)."""
        if getattr(self, "_fb", None) is None:
            fb = object.__new__(LightGlue)
            fb.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("_handle", "_fb", "_param", "_mx_cap", "_mx_dev", "_mx_pin")})
            fb.precision = "fp32x3"
            fb._handle = None
            fb._create(self.device)
            self._fb = fb
        self.range_fallbacks = getattr(self, "range_fallbacks", 0) + 1
        return self._fb

    @property
    def planes(self) -> int:
        """Operand planes of the engine actually in use: 1 bf16, 2 fp16x2, 3 bf16x3."""
        return int(lib.b2s_lg_planes(self._handle))

    def match_device(self, k0, d0, k1, d1, size0=None, size1=None, full=True):
        """CUDA f32 tensors k [m,2], d [m,128]; enqueues on the current stream WITHOUT any host
        synchronisation (early exit / pruning are decided on the device).  Returns a dict of device
        tensors; 'n' (valid rows of matches/scores) and 'stop' (executed layers) are device int32."""
        m, n = int(k0.shape[0]), int(k1.shape[0])
        dev = self.device
        k = max(min(m, n), 1)
        out = {"matches": torch.empty((k, 2), dtype=torch.int32, device=dev),
               "scores": torch.empty((k,), dtype=torch.float32, device=dev),
               "n": torch.zeros((1,), dtype=torch.int32, device=dev),
               "stop": torch.zeros((1,), dtype=torch.int32, device=dev)}
        ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
        if full:
            out.update(matches0=torch.empty((m,), dtype=torch.int32, device=dev), matches1=torch.empty((n,), dtype=torch.int32, device=dev),
                       matching_scores0=torch.empty((m,), dtype=torch.float32, device=dev),
                       matching_scores1=torch.empty((n,), dtype=torch.float32, device=dev),
                       prune0=torch.empty((m,), dtype=torch.int32, device=dev), prune1=torch.empty((n,), dtype=torch.int32, device=dev))
        s0 = (C.c_float * 2)(*[float(v) for v in size0]) if size0 is not None else None
        s1 = (C.c_float * 2)(*[float(v) for v in size1]) if size1 is not None else None
        st = torch.cuda.current_stream(dev).cuda_stream
        check(lib.b2s_lightglue_match(
            self._handle, k0.data_ptr(), d0.data_ptr(), m, k1.data_ptr(), d1.data_ptr(), n,
            C.cast(s0, C.c_void_p) if s0 is not None else None, C.cast(s1, C.c_void_p) if s1 is not None else None, st,
            out["matches"].data_ptr(), out["scores"].data_ptr(), out["n"].data_ptr(), out["stop"].data_ptr(),
            ptr(out.get("matches0")), ptr(out.get("matches1")), ptr(out.get("matching_scores0")),
            ptr(out.get("matching_scores1")), ptr(out.get("prune0")), ptr(out.get("prune1"))), "b2s_lightglue_match")
        return out

    @torch.no_grad()
    def __call__(self, data: dict) -> dict:
        """upstream LightGlue.forward contract (features_utils.py:157-161, :237)."""
        d0, d1 = data["image0"], data["image1"]
        with torch.cuda.device(self.device):
            prep = lambda t: t.to(self.device, dtype=torch.float32).contiguous()   # noqa: E731
            k0, k1 = prep(d0["keypoints"]), prep(d1["keypoints"])
            e0, e1 = prep(d0["descriptors"]), prep(d1["descriptors"])
            assert k0.dim() == 3 and k0.shape[0] == 1, "batch size must be 1"
            sz = []
            for d in (d0, d1):
                s = d.get("image_size")
                sz.append(None if s is None else [float(v) for v in torch.as_tensor(s).reshape(-1)[:2].tolist()])
            r = self.match_device(k0[0], e0[0], k1[0], e1[0], sz[0], sz[1], full=True)
            torch.cuda.current_stream(self.device).synchronize()
            nm = int(r["n"].item())
            if nm == _lib.LG_RANGE:
                return self.fallback()(data)
            return {"matches0": r["matches0"].long()[None], "matches1": r["matches1"].long()[None],
                    "matching_scores0": r["matching_scores0"][None], "matching_scores1": r["matching_scores1"][None],
                    "stop": int(r["stop"].item()), "matches": [r["matches"][:nm].long()], "scores": [r["scores"][:nm]],
                    "prune0": r["prune0"].long()[None], "prune1": r["prune1"].long()[None]}

    forward = __call__

    # -- batched matching: the pair is a grid dimension of every kernel (b2s_lightglue_match_batch_ex) ----------
    @property
    def max_batch(self) -> int:
        return int(lib.b2s_lg_max_batch())

    def reserve(self, max_kp: int, pairs: int = 1):
        """Allocate the workspace for `pairs` pairs per launch sequence up front (it otherwise grows on demand)."""
        check(lib.b2s_lg_reserve(self._handle, int(max_kp), int(pairs)), "b2s_lg_reserve")
        return self

    def match_batch_packed(self, kpts, desc, cu, pair_i, pair_j, stride=None, max_batch=0, out=None, counts=None, counts_dev=None):
        """Packed varlen form: kpts [T,2] / desc [T,128] CUDA f32 hold the features of F frames, frame f = rows
        [cu[f], cu[f+1]) - or [cu[f], cu[f] + counts[f]) when `counts` is given (frames in fixed-size slots) - with
        cu, counts, pair_i, pair_j host int arrays.  counts_dev (CUDA int32 [F], optional): the true count of frame f is
        min(counts[f], counts_dev[f]) read on the device (the extractor's device-resident n goes straight in; `counts` is then
        just the slot capacity and nothing waits for a count on the host).  Enqueues on the current stream without host
        synchronisation and returns device tensors {'matches' [P,stride,2] i32, 'scores' [P,stride] f32, 'n' [P] i32,
        'stop' [P] i32}; rows >= n[p] of pair p are undefined.  Results equal P single `match_device` calls."""
        cu = np.ascontiguousarray(cu, np.int32); pair_i = np.ascontiguousarray(pair_i, np.int32); pair_j = np.ascontiguousarray(pair_j, np.int32)
        P = len(pair_i)
        dev = self.device
        if counts is not None:
            counts = np.ascontiguousarray(counts, np.int32)
        n_frames = len(counts) if counts is not None else len(cu) - 1
        if stride is None:
            cnt = counts if counts is not None else np.diff(cu)
            stride = int(max(1, np.minimum(cnt[pair_i], cnt[pair_j]).max())) if P else 1
        if out is None:
            out = {"matches": torch.empty((max(P, 1), stride, 2), dtype=torch.int32, device=dev),
                   "scores": torch.empty((max(P, 1), stride), dtype=torch.float32, device=dev),
                   "n": torch.zeros((max(P, 1),), dtype=torch.int32, device=dev),
                   "stop": torch.zeros((max(P, 1),), dtype=torch.int32, device=dev)}
        st = torch.cuda.current_stream(dev).cuda_stream
        check(lib.b2s_lightglue_match_batch_ex(self._handle, kpts.data_ptr(), desc.data_ptr(), cu.ctypes.data,
                                               counts.ctypes.data if counts is not None else None,
                                               counts_dev.data_ptr() if counts_dev is not None else None, n_frames,
                                               pair_i.ctypes.data, pair_j.ctypes.data, P, st, int(stride), int(max_batch),
                                               out["matches"].data_ptr(), out["scores"].data_ptr(), out["n"].data_ptr(),
                                               out["stop"].data_ptr()), "b2s_lightglue_match_batch")
        return out

    def resolve_range(self, out, kpts, desc, cu, pair_i, pair_j, max_batch=0, counts=None, counts_dev=None) -> int:
        """Synchronises and re-runs, on the bf16x3 engine, the pairs of a `match_batch_packed` result that report
        n == LG_RANGE (precision='fp32' only; same arguments as that call).  Returns how many pairs were re-run."""
        n = out["n"].cpu().numpy()
        bad = np.nonzero(n[: len(pair_i)] == _lib.LG_RANGE)[0]
        if len(bad) == 0:
            return 0
        pi = np.ascontiguousarray(np.asarray(pair_i, np.int32)[bad]); pj = np.ascontiguousarray(np.asarray(pair_j, np.int32)[bad])
        r = self.fallback().match_batch_packed(kpts, desc, cu, pi, pj, int(out["matches"].shape[1]), max_batch, None, counts, counts_dev)
        idx = torch.as_tensor(bad, device=self.device, dtype=torch.long)
        for k in ("matches", "scores", "n", "stop"):
            out[k][idx] = r[k][: len(bad)]
        return int(len(bad))

    def match_batch_device(self, kpts_list, desc_list, pairs, stride=None, max_batch=0):
        """kpts_list[f] [n_f,2], desc_list[f] [n_f,128] CUDA f32; pairs = [(i, j), ...] frame indices."""
        with torch.cuda.device(self.device):
            cu = np.cumsum([0] + [int(k.shape[0]) for k in kpts_list]).astype(np.int32)
            kp = torch.cat([k.reshape(-1, 2) for k in kpts_list]).contiguous() if len(kpts_list) else torch.empty((0, 2), device=self.device)
            de = torch.cat([d.reshape(-1, 128) for d in desc_list]).contiguous() if len(desc_list) else torch.empty((0, 128), device=self.device)
            pi = np.asarray([p[0] for p in pairs], np.int32); pj = np.asarray([p[1] for p in pairs], np.int32)
            out = self.match_batch_packed(kp, de, cu, pi, pj, stride, max_batch)
            out["_keepalive"] = (kp, de)       # the packed inputs must outlive the enqueued kernels
            if self.precision == "fp32":
                self.resolve_range(out, kp, de, cu, pi, pj, max_batch)
            return out

    def match_mixed(self, k0, d0, k1, d1):
        """The drop-in's matching call: each side is either host numpy (k [m,2], d [m,128] f32: uploaded) or CUDA tensors
        already on this device (a frame whose features the feature cache kept on the GPU: no upload).  One packed D2H of
        (count, executed layers, matches, scores) through pinned memory, one synchronisation.  Returns
        {'matches' int32 [K,2], 'scores' f32 [K], 'stop', 'h2d_bytes'}."""
        dev = self.device
        with torch.cuda.device(dev):
            h2d = 0
            ts = []
            for a in (k0, d0, k1, d1):
                if isinstance(a, torch.Tensor):
                    ts.append(a)
                else:
                    a = np.ascontiguousarray(a, np.float32)
                    h2d += a.nbytes
                    ts.append(torch.from_numpy(a).to(dev, non_blocking=True))
            k0, d0, k1, d1 = ts
            m, n = int(k0.shape[0]), int(k1.shape[0])
            cap = max(min(m, n), 1)
            if getattr(self, "_mx_cap", 0) < cap:
                self._mx_cap = max(cap, 2048)
                self._mx_dev = torch.zeros(4 + 3 * self._mx_cap, dtype=torch.int32, device=dev)     # n, stop, -, - | matches | scores
                self._mx_pin = torch.zeros(4 + 3 * self._mx_cap, dtype=torch.int32).pin_memory()
            o, c = self._mx_dev, self._mx_cap
            base = o.data_ptr()
            st = torch.cuda.current_stream(dev)
            check(lib.b2s_lightglue_match(self._handle, k0.data_ptr(), d0.data_ptr(), m, k1.data_ptr(), d1.data_ptr(), n, None, None,
                                          st.cuda_stream, base + 16, base + 16 + 8 * c, base, base + 4, None, None, None, None, None, None),
                  "b2s_lightglue_match")
            self._mx_pin.copy_(o, non_blocking=True)
            st.synchronize()
            hp = self._mx_pin.numpy()
            nm, stop = int(hp[0]), int(hp[1])
            if nm == _lib.LG_RANGE:
                r = self.fallback().match_mixed(k0, d0, k1, d1)
                r["h2d_bytes"] += h2d
                return r
            matches = hp[4: 4 + 2 * nm].reshape(nm, 2).copy()
            scores = hp[4 + 2 * c: 4 + 2 * c + nm].view(np.float32).copy()
            return {"matches": matches, "scores": scores, "stop": stop, "h2d_bytes": h2d}

    def match_host(self, k0, d0, k1, d1, size0=None, size1=None, full=False):
        """One C-ABI call with host (numpy f32) buffers: b2s_lightglue_match_host."""
        k0 = np.ascontiguousarray(k0, np.float32); k1 = np.ascontiguousarray(k1, np.float32)
        d0 = np.ascontiguousarray(d0, np.float32); d1 = np.ascontiguousarray(d1, np.float32)
        m, n = len(k0), len(k1)
        k = max(min(m, n), 1)
        matches = np.empty((k, 2), np.int32); scores = np.empty((k,), np.float32)
        nm, stop = C.c_int32(0), C.c_int32(0)
        extra = {}
        if full:
            extra = dict(matches0=np.empty(m, np.int32), matches1=np.empty(n, np.int32), matching_scores0=np.empty(m, np.float32),
                         matching_scores1=np.empty(n, np.float32), prune0=np.empty(m, np.int32), prune1=np.empty(n, np.int32))
        p = lambda a: a.ctypes.data if a is not None else None   # noqa: E731
        s0 = np.asarray(size0, np.float32) if size0 is not None else None
        s1 = np.asarray(size1, np.float32) if size1 is not None else None
        check(lib.b2s_lightglue_match_host(
            self._handle, p(k0), p(d0), m, p(k1), p(d1), n, p(s0), p(s1), p(matches), p(scores), C.addressof(nm),
            C.addressof(stop), p(extra.get("matches0")), p(extra.get("matches1")), p(extra.get("matching_scores0")),
            p(extra.get("matching_scores1")), p(extra.get("prune0")), p(extra.get("prune1"))), "b2s_lightglue_match_host")
        out = {"matches": matches[:nm.value], "scores": scores[:nm.value], "stop": stop.value}
        out.update(extra)
        return out

    def profile(self, on=True):
        check(lib.b2s_lg_profile(self._handle, 1 if on else 0), "b2s_lg_profile")

    def workspace_bytes(self, max_kp: int, pairs: int = 1) -> int:
        cfg = _lib.LgCfg()
        lib.b2s_lg_default_cfg(C.byref(cfg))
        cfg.n_layers, cfg.precision = self.n_layers, _lib.PRECISIONS[self.precision]
        return int(lib.b2s_lg_workspace_bytes(C.byref(cfg), int(max_kp), int(pairs)))

    def profile_read(self, cls: int):
        """(summed kernel ms, launches) of class 0 = attention, 1 = GEMM since the last read."""
        ms, n = C.c_double(0), C.c_longlong(0)
        check(lib.b2s_lg_profile_read(self._handle, cls, C.byref(ms), C.byref(n)), "b2s_lg_profile_read")
        return ms.value, n.value

    def profile_work(self):
        """(sum nq*nk over executed self-attention problems, same for cross-attention) since profile(True)."""
        a, b = C.c_double(0), C.c_double(0)
        check(lib.b2s_lg_profile_work(self._handle, C.byref(a), C.byref(b)), "b2s_lg_profile_work")
        return a.value, b.value

    def set_debug(self, on=True):
        check(lib.b2s_lg_set_debug(self._handle, 1 if on else 0), "b2s_lg_set_debug")

    def debug(self, name: str) -> np.ndarray:
        n = C.c_size_t(0)
        check(lib.b2s_lg_debug_get(self._handle, name.encode(), None, 0, C.byref(n)), "debug_get")
        out = np.empty((max(n.value, 1),), np.float32)
        check(lib.b2s_lg_debug_get(self._handle, name.encode(), out.ctypes.data, n.value, C.byref(n)), "debug_get")
        return out[:n.value]

    @property
    def launches(self) -> int:
        return int(lib.b2s_lg_launch_count(self._handle))
