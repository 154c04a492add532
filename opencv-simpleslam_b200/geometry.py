"""Host side of the rows directly behind the matcher, all on libb200slam.so:

* `FundamentalRansac` / `find_fundamental_mat` - stands in for
  `cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, 0.99)` as called by `filter_matches_ransac`
  (/root/reference/slam/core/features_utils.py:185-200);
* `FrameUndistorter` - stands in for `cv2.remap(img, mapx, mapy, cv2.INTER_LINEAR)` in the frame loop
  (/root/reference/slam/monocular/main_revamped.py:313-315,323-324), bit-exact with OpenCV on u8 BGR frames.

No CPU fallback: constructing either object without a CUDA device raises."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import lib, check

DEFAULT_HYPOTHESES = 2048      # cv2's sequential loop draws at most 2000 samples


def _device_index(device=None) -> int:
    if not torch.cuda.is_available():
        raise RuntimeError("b200slam needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.cuda.current_device()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"b200slam runs on CUDA only, got device {dev}")
    return dev.index if dev.index is not None else torch.cuda.current_device()


class FundamentalRansac:
    """Parallel 7-point RANSAC (b2s_fm_*).  One handle per device; grows on demand."""

    def __init__(self, device=None, max_points: int = 4096, n_hyp: int = DEFAULT_HYPOTHESES, seed: int = 0):
        self.device_index, self.n_hyp, self.seed = _device_index(device), int(n_hyp), int(seed)
        self.max_hyp = max(self.n_hyp, 1000)       # handle capacity: the parallel estimator's samples / cv2's default maxIters
        self._handle, self.max_points = None, 0
        self._create(max(int(max_points), 8))

    def _create(self, max_points: int):
        self._destroy()
        h = C.c_void_p()
        check(lib.b2s_fm_create(self.device_index, max_points, self.max_hyp, C.byref(h)), "b2s_fm_create")
        self._handle, self.max_points = h, max_points

    def _destroy(self):
        if getattr(self, "_handle", None):
            lib.b2s_fm_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def run_host(self, pts1: np.ndarray, pts2: np.ndarray, thresh: float = 1.0, seed: int | None = None):
        """numpy [n,2] correspondences -> (F float64 [3,3] or None, mask uint8 [n,1] or None) like cv2.findFundamentalMat.
        Also leaves `last_count` / `last_model_index` (sample*3 + root) on the object."""
        pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
        n = len(pts1)
        if len(pts2) != n:
            raise ValueError("pts1 and pts2 must have the same length")
        if n < 7:                                   # cv2 returns (None, None) below the minimal sample size
            self.last_count, self.last_model_index = 0, -1
            return None, None
        if n > self.max_points:
            self._create(int(2 ** np.ceil(np.log2(n))))
        mask = np.empty((n,), np.uint8)
        F = np.empty((9,), np.float64)
        cnt, idx = C.c_int32(0), C.c_int32(0)
        check(lib.b2s_fm_ransac_host(self._handle, pts1.ctypes.data, pts2.ctypes.data, n, float(thresh), self.n_hyp,
                                     C.c_uint64(self.seed if seed is None else int(seed)), mask.ctypes.data, F.ctypes.data,
                                     C.addressof(cnt), C.addressof(idx)), "b2s_fm_ransac_host")
        self.last_count, self.last_model_index = cnt.value, idx.value
        if idx.value < 0:
            return None, None
        return F.reshape(3, 3), mask.reshape(-1, 1)

    def run_cv(self, pts1: np.ndarray, pts2: np.ndarray, thresh: float = 3.0, confidence: float = 0.99, max_iters: int = 1000):
        """The cv2-IDENTICAL estimator (b2s_fm_cv_ransac_host): same (F, mask) as
        `cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, confidence)` for n >= 15 correspondences.
        Leaves `last_count` and `last_info` = (winning iteration, model, final iteration bound, subsets drawn)."""
        pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
        n = len(pts1)
        if len(pts2) != n:
            raise ValueError("pts1 and pts2 must have the same length")
        if n < 15:
            raise ValueError("run_cv covers OpenCV's RANSAC branch (n >= 15); below that cv2 runs LMedS")
        if n > self.max_points or max_iters > self.max_hyp:
            self.max_hyp = max(self.max_hyp, int(max_iters))
            self._create(int(2 ** np.ceil(np.log2(max(n, self.max_points)))))
        mask = np.empty((n,), np.uint8)
        F = np.empty((9,), np.float64)
        cnt = C.c_int32(0)
        info = (C.c_int32 * 4)()
        check(lib.b2s_fm_cv_ransac_host(self._handle, pts1.ctypes.data, pts2.ctypes.data, n, float(thresh), float(confidence),
                                        int(max_iters), mask.ctypes.data, F.ctypes.data, C.addressof(cnt), info), "b2s_fm_cv_ransac_host")
        self.last_count, self.last_info = cnt.value, tuple(info)
        if info[0] < 0:
            return None, None
        return F.reshape(3, 3), mask.reshape(-1, 1)

    def run_device(self, kp0: torch.Tensor, kp1: torch.Tensor, pairs: torch.Tensor | None, n: int, thresh: float = 1.0,
                   seed: int | None = None):
        """Device-resident form: kp0/kp1 CUDA f32 [*,2], pairs CUDA int32 [>=n,2] (the matcher's `matches`) or None.
        Enqueues on the current stream, no host sync.  Returns device tensors (mask u8 [n], F f64 [9], result i32 [2])."""
        if n > self.max_points:
            self._create(int(2 ** np.ceil(np.log2(n))))
        dev = kp0.device
        mask = torch.empty((n,), dtype=torch.uint8, device=dev)
        F = torch.empty((9,), dtype=torch.float64, device=dev)
        res = torch.empty((2,), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        check(lib.b2s_fm_ransac(self._handle, kp0.data_ptr(), kp1.data_ptr(), pairs.data_ptr() if pairs is not None else None,
                                int(n), float(thresh), self.n_hyp, C.c_uint64(self.seed if seed is None else int(seed)), st,
                                mask.data_ptr(), F.data_ptr(), res.data_ptr()), "b2s_fm_ransac")
        return mask, F, res

    def debug_models(self, hyp: int):
        m = np.empty((27,), np.float64); n = C.c_int32(0); c = np.empty((3,), np.int32)
        check(lib.b2s_fm_debug_models(self._handle, int(hyp), m.ctypes.data, C.addressof(n), c.ctypes.data), "b2s_fm_debug_models")
        return m.reshape(3, 3, 3)[:n.value], c[:n.value]

    @property
    def launches(self) -> int:
        return int(lib.b2s_fm_launch_count(self._handle))


_default_ransac: dict = {}


def default_ransac() -> FundamentalRansac:
    """Process-wide handle of the current device (filter_matches_ransac has no handle argument in the reference)."""
    idx = _device_index()
    if idx not in _default_ransac:
        _default_ransac[idx] = FundamentalRansac(idx)
    return _default_ransac[idx]


def find_fundamental_mat(pts1, pts2, thresh: float = 1.0):
    """(F, mask) with cv2.findFundamentalMat's return convention."""
    return default_ransac().run_host(pts1, pts2, thresh)


def find_fundamental_mat_cv(pts1, pts2, thresh: float = 3.0, confidence: float = 0.99):
    """`cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, confidence)` with identical results: OpenCV's RANSAC
    (n >= 15) on the GPU, the degenerate small cases (n < 15: OpenCV's LMedS branch, whose winner for n <= 13 is decided
    by the rounding noise of its own arithmetic; n == 7; n < 7) by cv2 itself.  One documented difference: when no valid
    7-point sample exists (all points collinear / identical) cv2 returns F = None together with an UNINITIALISED mask
    array; this function returns (None, None), i.e. `filter_matches_ransac` keeps no match."""
    pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    if len(pts1) < 15:
        import cv2
        return cv2.findFundamentalMat(pts1, np.ascontiguousarray(pts2, np.float32).reshape(-1, 2), cv2.FM_RANSAC, thresh, confidence)
    return default_ransac().run_cv(pts1, pts2, thresh, confidence)


class FrameUndistorter:
    """cv2.remap with fixed maps on the GPU (b2s_remap_*).  `remap(img)` returns a host image like cv2.remap;
    `remap_to_device(img)` leaves the undistorted u8 frame on the device for `ALIKED.extract_device`."""

    def __init__(self, mapx: np.ndarray, mapy: np.ndarray, src_shape=None, device=None):
        self.mapx = np.ascontiguousarray(mapx, np.float32)
        self.mapy = np.ascontiguousarray(mapy, np.float32)
        if self.mapx.ndim != 2 or self.mapx.shape != self.mapy.shape:
            raise ValueError("mapx / mapy must be float32 [H,W] maps of equal shape (cv2.CV_32FC1)")
        self.dH, self.dW = self.mapx.shape
        self.sH, self.sW = (int(src_shape[0]), int(src_shape[1])) if src_shape is not None else (self.dH, self.dW)
        self.device_index = _device_index(device)
        h = C.c_void_p()
        check(lib.b2s_remap_create(self.device_index, self.mapx.ctypes.data, self.mapy.ctypes.data, self.dH, self.dW,
                                   self.sH, self.sW, C.byref(h)), "b2s_remap_create")
        self._handle = h

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                lib.b2s_remap_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _check(self, img):
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3 or img.shape[:2] != (self.sH, self.sW):
            raise ValueError(f"expected a {self.sH}x{self.sW}x3 uint8 BGR frame, got {img.shape} {img.dtype}")
        return np.ascontiguousarray(img)

    def remap(self, img: np.ndarray) -> np.ndarray:
        img = self._check(img)
        out = np.empty((self.dH, self.dW, 3), np.uint8)
        check(lib.b2s_remap_bgr_host(self._handle, img.ctypes.data, 3 * self.sW, out.ctypes.data, 3 * self.dW), "b2s_remap_bgr_host")
        return out

    def remap_to_device(self, img: np.ndarray) -> torch.Tensor:
        """Upload + remap on the current stream; returns a CUDA u8 [H,W,3] tensor (fresh memory)."""
        img = self._check(img)
        dev = torch.device("cuda", self.device_index)
        with torch.cuda.device(dev):
            src = torch.from_numpy(img).to(dev, non_blocking=True)
            dst = torch.empty((self.dH, self.dW, 3), dtype=torch.uint8, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            check(lib.b2s_remap_bgr(self._handle, src.data_ptr(), 3 * self.sW, st, dst.data_ptr(), 3 * self.dW), "b2s_remap_bgr")
            src.record_stream(torch.cuda.current_stream(dev))
        return dst

    @property
    def launches(self) -> int:
        return int(lib.b2s_remap_launch_count(self._handle))
