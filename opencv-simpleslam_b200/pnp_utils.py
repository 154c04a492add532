"""Drop-in for the landmark re-association step of `/root/reference/slam/core/pnp_utils.py`:
`reproject_and_match_2d3d` (:224-304) with the same signature and `Matches2D3D` return type, for the float descriptors
the ALIKED frontend produces.  Projection, window search and descriptor distances run in libb200slam.so
(b2s_reproj_match); the host keeps a device-resident mirror of the map's most recent descriptors so that only
landmarks whose observation list changed since the previous frame are re-uploaded.

Reference quirks kept: `use_cosine` does not change the distance (the reference always evaluates L2 for float
descriptors, pnp_utils.py:121) and `max_hamm` is unused for float descriptors."""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from ._lib import lib, check
from .features_utils import _kps_to_array
from .geometry import _device_index

MAX_OBS = 6            # the reference checks the last six observations (pnp_utils.py:112)
CAND_CAP = 64          # candidates kept per landmark (only keypoints passing the descriptor gate are stored, nearest first)
CAND_CAP_MAX = 256     # library limit (REPROJ_MAXCAP): used for the one retry when a truncated list ran dry
MAX_KPS_LIB = 49152    # library limit on keypoints per frame (REPROJ_MAX_KPS)


@dataclass
class Matches2D3D:     # pnp_utils.py:51-56
    pts3d: np.ndarray          # (N,3) world
    pts2d: np.ndarray          # (N,2) image
    kp_indices: List[int]      # indices into current frame keypoints
    mp_ids: List[int]          # matched map point ids


def _kp_coords(kps) -> np.ndarray:
    """list[cv2.KeyPoint] / KeyPointArray / ndarray (N,>=2) -> float32 (N,2)   (pnp_utils.py:65-76)"""
    if isinstance(kps, np.ndarray):
        if kps.size == 0:
            return np.empty((0, 2), np.float32)
        assert kps.ndim == 2 and kps.shape[1] >= 2, "kps must be (N,2)"
        return np.ascontiguousarray(kps[:, :2], dtype=np.float32)
    return _kps_to_array(kps)


def _empty() -> Matches2D3D:
    return Matches2D3D(np.zeros((0, 3), np.float32), np.zeros((0, 2), np.float32), [], [])


class MapDescriptorMirror:
    """Device table [rows, 6, 128] of every landmark's most recent descriptors + usable count, keyed by landmark id.
    `sync(world_map)` walks the map once (O(P) Python, no per-descriptor work for unchanged landmarks), uploads the
    rows whose observation list changed and returns (ids, positions f64 [P,3], rows i32 [P]) in map order."""

    def __init__(self, device_index: int, capacity: int = 4096):
        self.dev = torch.device("cuda", device_index)
        self.row_of: dict = {}
        self.sig: dict = {}
        self.free: list = []
        self.capacity = 0
        self.desc = self.nobs = None
        self._grow(capacity)

    def _grow(self, capacity: int):
        desc = torch.zeros((capacity, MAX_OBS, 128), dtype=torch.float32, device=self.dev)
        nobs = torch.zeros((capacity,), dtype=torch.int32, device=self.dev)
        if self.desc is not None:
            desc[:self.capacity] = self.desc; nobs[:self.capacity] = self.nobs
        self.free.extend(range(capacity - 1, self.capacity - 1, -1))
        self.desc, self.nobs, self.capacity = desc, nobs, capacity

    @staticmethod
    def _pack(observations):
        """(usable count, [count,128] f32): non-None descriptors among the last six, oldest first; count 0 if the
        newest observation has no descriptor (the reference then skips the landmark, pnp_utils.py:276-278)."""
        if not observations or observations[-1][2] is None:
            return 0, None
        ds = [np.asarray(d, np.float32).reshape(-1) for _, _, d in observations[-MAX_OBS:] if d is not None]
        if any(d.dtype == np.uint8 for _, _, d in observations[-MAX_OBS:] if d is not None and hasattr(d, "dtype")):
            raise NotImplementedError("binary (uint8) descriptors are outside the ALIKED frontend; use the reference's OpenCV path")
        ds = [d for d in ds if d.shape[0] == 128]
        return len(ds), (np.stack(ds) if ds else None)

    def sync(self, world_map):
        items = world_map.points
        P = len(items)
        ids = list(items.keys())
        rows = np.empty((P,), np.int32)
        pos = np.empty((P, 3), np.float64)
        upd_rows, upd_desc, upd_n = [], [], []
        if len(self.row_of) >= P:                              # landmarks culled from the map release their rows
            for k in [k for k in self.row_of if k not in items]:
                self.free.append(self.row_of.pop(k)); self.sig.pop(k, None)
        for i, (pid, mp) in enumerate(items.items()):
            pos[i] = mp.position
            obs = mp.observations
            row = self.row_of.get(pid)
            if row is None:
                if not self.free:
                    self._grow(self.capacity * 2)
                row = self.free.pop()
                self.row_of[pid] = row
            rows[i] = row
            sig = (len(obs), id(obs[-1][2]) if obs else 0)
            if self.sig.get(pid) != sig:
                self.sig[pid] = sig
                n, d = self._pack(obs)
                block = np.zeros((MAX_OBS, 128), np.float32)
                if n:
                    block[:n] = d
                upd_rows.append(row); upd_desc.append(block); upd_n.append(n)
        if upd_rows:
            r = torch.from_numpy(np.asarray(upd_rows, np.int64)).to(self.dev)
            self.desc.index_copy_(0, r, torch.from_numpy(np.stack(upd_desc)).to(self.dev))
            self.nobs.index_copy_(0, r, torch.from_numpy(np.asarray(upd_n, np.int32)).to(self.dev))
        return ids, pos, rows


class ReprojectionMatcher:
    """b2s_reproj_* handle + per-map descriptor mirrors."""

    def __init__(self, device=None, max_points: int = 8192, max_kps: int = 4096, cand_cap: int = CAND_CAP):
        self.device_index = _device_index(device)
        self._handle, self.max_points, self.max_kps, self.cand_cap = None, 0, 0, int(cand_cap)
        self._wide = None                      # lazily created twin with CAND_CAP_MAX candidates per landmark
        self._create(max_points, max_kps)
        self._mirrors = weakref.WeakKeyDictionary()

    def _create(self, max_points, max_kps):
        if max_kps > MAX_KPS_LIB:
            raise ValueError(f"reproject_and_match_2d3d: {max_kps} keypoints exceed the library limit {MAX_KPS_LIB}")
        h = C.c_void_p()
        # create the new handle first: if that fails the old one stays valid (and is destroyed exactly once)
        check(lib.b2s_reproj_create(self.device_index, int(max_points), int(max_kps), self.cand_cap, C.byref(h)), "b2s_reproj_create")
        old, self._handle = self._handle, h
        self.max_points, self.max_kps = int(max_points), int(max_kps)
        if old:
            lib.b2s_reproj_destroy(old)

    def __del__(self):
        try:
            h, self._handle = getattr(self, "_handle", None), None
            if h:
                lib.b2s_reproj_destroy(h)
        except Exception:
            pass

    def mirror_of(self, world_map) -> MapDescriptorMirror:
        try:
            m = self._mirrors.get(world_map)
            if m is None:
                m = self._mirrors[world_map] = MapDescriptorMirror(self.device_index)
            return m
        except TypeError:                                       # map object not weak-referenceable: no caching
            return MapDescriptorMirror(self.device_index)

    def match(self, world_map, K, Tcw_pred, kps_cur, des_cur, img_w, img_h, radius_px=12.0, max_l2=0.8) -> Matches2D3D:
        pts2d_cur = _kp_coords(kps_cur)
        N, P = len(pts2d_cur), len(world_map.points)
        if N == 0 or P == 0:
            return _empty()
        des = np.ascontiguousarray(des_cur, np.float32)
        if des.shape != (N, 128):
            raise ValueError(f"expected float descriptors [N,128] matching the keypoints, got {des.shape}")
        if P > self.max_points or N > self.max_kps:
            grow = lambda need, lim: min(int(2 ** np.ceil(np.log2(need))), lim) if lim else int(2 ** np.ceil(np.log2(need)))   # noqa: E731
            self._create(max(self.max_points, grow(P, 0)), max(self.max_kps, max(grow(N, MAX_KPS_LIB), min(N, MAX_KPS_LIB))))
        dev = torch.device("cuda", self.device_index)
        with torch.cuda.device(dev):
            mir = self.mirror_of(world_map)
            ids, pos, rows = mir.sync(world_map)
            Xw = torch.from_numpy(pos).to(dev, non_blocking=True)
            rows_d = torch.from_numpy(rows).to(dev, non_blocking=True)
            kps_d = torch.from_numpy(pts2d_cur).to(dev, non_blocking=True)
            des_d = torch.from_numpy(des).to(dev, non_blocking=True)
            out = torch.empty((P,), dtype=torch.int32, device=dev)
            flags = torch.zeros((1,), dtype=torch.int32, device=dev)
            Kh = np.ascontiguousarray(K, np.float64).reshape(9)
            Th = np.ascontiguousarray(Tcw_pred, np.float64).reshape(16)
            st = torch.cuda.current_stream(dev).cuda_stream
            check(lib.b2s_reproj_match(self._handle, Xw.data_ptr(), mir.desc.data_ptr(), mir.nobs.data_ptr(), rows_d.data_ptr(), P,
                                       MAX_OBS, Kh.ctypes.data, Th.ctypes.data, kps_d.data_ptr(), des_d.data_ptr(), N,
                                       int(img_w), int(img_h), float(radius_px), float(max_l2), st, out.data_ptr(), None,
                                       flags.data_ptr()), "b2s_reproj_match")
            res = torch.cat([out, flags]).cpu().numpy()
        if res[-1] & 1:
            # a landmark had more than cand_cap keypoints passing the descriptor gate AND lost all of the kept (nearest)
            # ones to earlier landmarks: rerun once with the library's widest candidate list before giving up
            if self.cand_cap < CAND_CAP_MAX:
                if self._wide is None:
                    self._wide = ReprojectionMatcher(self.device_index, self.max_points, self.max_kps, CAND_CAP_MAX)
                    self._wide._mirrors = self._mirrors
                return self._wide.match(world_map, K, Tcw_pred, kps_cur, des_cur, img_w, img_h, radius_px, max_l2)
            raise RuntimeError(f"reproject_and_match_2d3d: a landmark exhausted {self.cand_cap} descriptor-gated candidates "
                               f"(radius_px={radius_px}, max_l2={max_l2})")
        kp_of = res[:P]
        sel = np.flatnonzero(kp_of >= 0)
        if len(sel) == 0:
            return _empty()
        kpids = kp_of[sel].astype(np.int64)
        return Matches2D3D(pos[sel].astype(np.float32), pts2d_cur[kpids], kpids.tolist(), [ids[i] for i in sel.tolist()])

    @property
    def launches(self) -> int:
        return int(lib.b2s_reproj_launch_count(self._handle))


_default: dict = {}


def default_matcher() -> ReprojectionMatcher:
    idx = _device_index()
    if idx not in _default:
        _default[idx] = ReprojectionMatcher(idx)
    return _default[idx]


def reproject_and_match_2d3d(world_map, K, Tcw_pred, kps_cur, des_cur, img_w: int, img_h: int, radius_px: float = 12.0,
                             max_hamm: int = 64, max_l2: float = 0.8, use_cosine: bool = False) -> Matches2D3D:
    """Same signature and result as the reference (pnp_utils.py:224-304) for float descriptors."""
    if des_cur is None or len(des_cur) == 0:
        return _empty()
    if not world_map.points:
        return _empty()
    if getattr(des_cur, "dtype", None) == np.uint8:
        raise NotImplementedError("binary (uint8) descriptors are outside the ALIKED frontend; use the reference's OpenCV path")
    return default_matcher().match(world_map, K, Tcw_pred, kps_cur, des_cur, img_w, img_h, radius_px, max_l2)
