"""TEST INFRASTRUCTURE ONLY (never imported by the product package).

numpy restatement of `reproject_and_match_2d3d` for float descriptors
(/root/reference/slam/core/pnp_utils.py:224-304 with its helpers `_project_points` :126-145, `_kp_coords` :65-76,
`_choose_mp_descriptor` :45-49, `_best_mp_distance_to_cur_desc` :112-123, `_desc_distance` :78-109).
Differences from the reference's text: the cKDTree ball query is a brute-force `d^2 <= r^2` test (same inclusive
bound, float64 on float32 inputs as the tree does), and candidates are visited in ascending keypoint index (the tree's
visiting order only matters for exact distance ties).  Pinned against the reference function itself, imported from
/root/reference in the build container: tests/golden/make_golden_pnp.py -> tests/golden/pnp_reproj.npz."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class Matches2D3D:                      # pnp_utils.py:51-56
    pts3d: np.ndarray
    pts2d: np.ndarray
    kp_indices: List[int]
    mp_ids: List[int]


def project_points(K, Tcw, pts_w):      # pnp_utils.py:126-145
    pts_w = np.asarray(pts_w, dtype=np.float64)
    Xc = pts_w @ Tcw[:3, :3].T + Tcw[:3, 3]
    z = Xc[:, 2]
    uv = np.full((len(pts_w), 2), -1.0, dtype=np.float32)
    valid = z > 1e-8
    if np.any(valid):
        proj = (K @ (Xc[valid] / z[valid, None]).T).T
        uv[valid] = proj[:, :2].astype(np.float32)
    return uv, z


def desc_distance(a, b, metric):        # pnp_utils.py:78-109 (float branch)
    a = np.asarray(a, np.float32).reshape(-1); b = np.asarray(b, np.float32).reshape(-1)
    if a.shape != b.shape:
        return float("inf")
    if metric == "cosine":
        na, nb = float(np.linalg.norm(a)), float(np.linalg.norm(b))
        if na == 0.0 or nb == 0.0:
            return float("inf")
        return 1.0 - float(np.dot(a, b) / (na * nb))
    return float(np.linalg.norm(a - b))


def reproject_and_match_2d3d(world_map, K, Tcw_pred, kps_cur, des_cur, img_w, img_h, radius_px=12.0, max_l2=0.8,
                             use_cosine=False, max_obs_check=6):
    empty = Matches2D3D(np.zeros((0, 3), np.float32), np.zeros((0, 2), np.float32), [], [])
    if des_cur is None or len(des_cur) == 0 or not world_map.points:
        return empty
    if isinstance(kps_cur, (list, tuple)) and len(kps_cur) and hasattr(kps_cur[0], "pt"):
        pts2d_cur = np.float32([kp.pt for kp in kps_cur])
    else:
        pts2d_cur = np.asarray(kps_cur, np.float32).reshape(-1, 2)
    if len(pts2d_cur) == 0:
        return empty
    items = list(world_map.points.items())
    pts3d_all = np.asarray([mp.position for _, mp in items], dtype=np.float64)
    uv_all, z_all = project_points(np.asarray(K, np.float64), np.asarray(Tcw_pred, np.float64), pts3d_all)
    cand = (z_all > 0.0) & (uv_all[:, 0] >= 0.0) & (uv_all[:, 0] < float(img_w)) & (uv_all[:, 1] >= 0.0) & (uv_all[:, 1] < float(img_h))
    # reference quirk: `use_cosine` only selects which threshold variable is read (max_l2 either way, :262-263); the
    # distance itself always goes through _desc_distance(metric="auto") = L2 for float descriptors (:121)
    metric = "l2"
    used = np.zeros(len(pts2d_cur), bool)
    kp64 = pts2d_cur.astype(np.float64)
    pts3d, pts2d, kpids, mpids = [], [], [], []
    for arr_idx in np.flatnonzero(cand):
        mp_id, mp = items[arr_idx]
        d2 = ((kp64 - uv_all[arr_idx].astype(np.float64)) ** 2).sum(axis=1)
        cand_idx = np.flatnonzero(d2 <= float(radius_px) ** 2)
        if len(cand_idx) == 0:
            continue
        if not mp.observations or mp.observations[-1][2] is None:
            continue
        best_i, best_d = -1, 1e9
        for i in cand_idx:
            if used[i]:
                continue
            d = float("inf")
            for _, _, dd in mp.observations[-max_obs_check:]:
                if dd is not None:
                    d = min(d, desc_distance(dd, des_cur[i], metric))
            if d < best_d:
                best_d, best_i = d, int(i)
        if best_i < 0 or best_d > max_l2:
            continue
        used[best_i] = True
        pts3d.append(pts3d_all[arr_idx].astype(np.float32)); pts2d.append(pts2d_cur[best_i])
        kpids.append(best_i); mpids.append(mp_id)
    if not pts3d:
        return empty
    return Matches2D3D(np.asarray(pts3d, np.float32), np.asarray(pts2d, np.float32), kpids, mpids)


# ---- seeded scene shared by the golden generator and the tests --------------------------------------------------
@dataclass
class _MP:
    id: int
    position: np.ndarray
    observations: list


class _Map:
    def __init__(self):
        self.points = {}


def tracking_scene(n_points=1500, n_kps=2048, seed=0, W=1241, H=376, desc_noise=0.25, crowd=True):
    """A map of landmarks in front of a KITTI-like camera, a predicted pose, and a current frame whose keypoints are
    noisy projections of a subset (so windows hold several candidates and keypoints are contested).  Returns
    (world_map, K, Tcw, kps f32 [N,2], des f32 [N,128])."""
    rng = np.random.default_rng(seed)
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.2], [0, 0, 1.0]])
    ang = 0.02
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    Tcw = np.eye(4); Tcw[:3, :3] = R; Tcw[:3, 3] = [0.1, -0.05, -0.6]
    X = np.stack([rng.uniform(-25, 25, n_points), rng.uniform(-4, 4, n_points), rng.uniform(-5, 70, n_points)], axis=1)
    unit = lambda v: (v / (np.linalg.norm(v, axis=-1, keepdims=True) + 1e-8)).astype(np.float32)   # noqa: E731
    base = unit(rng.normal(size=(n_points, 128)))
    wm = _Map()
    for i in range(n_points):
        n_obs = int(rng.integers(0, 9))                    # 0..8 observations: exercises the last-6 window and empties
        obs = []
        for o in range(n_obs):
            d = unit(base[i] + desc_noise * 0.5 * rng.normal(size=128) / np.sqrt(128))
            obs.append((int(rng.integers(0, 50)), int(rng.integers(0, n_kps)), d))
        if n_obs and rng.random() < 0.05:
            obs[-1] = (obs[-1][0], obs[-1][1], None)       # newest descriptor missing -> the reference skips the point
        if n_obs > 2 and rng.random() < 0.1:
            obs[-2] = (obs[-2][0], obs[-2][1], None)       # a missing one inside the window is just ignored
        wm.points[1000 + 3 * i] = _MP(1000 + 3 * i, X[i].copy(), obs)
    Xc = X @ R.T + Tcw[:3, 3]
    uv = (K @ (Xc / Xc[:, 2:3]).T).T[:, :2]
    vis = np.flatnonzero((Xc[:, 2] > 0.5) & (uv[:, 0] >= 0) & (uv[:, 0] < W) & (uv[:, 1] >= 0) & (uv[:, 1] < H))
    kps = np.stack([rng.uniform(0, W, n_kps), rng.uniform(0, H, n_kps)], axis=1)
    des = unit(rng.normal(size=(n_kps, 128)))
    take = rng.permutation(vis)[:min(len(vis), n_kps // 2)]
    for slot, i in enumerate(take):
        kps[slot] = uv[i] + rng.normal(0, 3.0, 2)
        des[slot] = unit(base[i] + desc_noise * rng.normal(size=128) / np.sqrt(128))
    if crowd:                                              # extra keypoints near projections: contested windows
        extra = rng.permutation(vis)[:n_kps // 4]
        for slot, i in enumerate(extra, start=n_kps // 2):
            kps[slot] = uv[i] + rng.normal(0, 5.0, 2)
            des[slot] = unit(base[i] + 1.5 * desc_noise * rng.normal(size=128) / np.sqrt(128))
    return wm, K, Tcw, kps.astype(np.float32), des.astype(np.float32)
