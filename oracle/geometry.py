"""TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatements for the rows behind the matcher:

* `remap_bgr_u8`   - OpenCV's fixed-point bilinear `cv2.remap(src_u8, mapx_f32, mapy_f32, INTER_LINEAR)` as the
                     reference calls it at slam/monocular/main_revamped.py:313-315,323-324.  Pinned bit-exactly
                     against cv2.remap itself (tests/test_oracle_pins.py) - cv2 IS the reference's implementation.
* `fm_ransac`      - the parallel 7-point fundamental-matrix RANSAC that stands in for
                     `cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, 0.99)` in `filter_matches_ransac`
                     (slam/core/features_utils.py:185-200).  cv2's sampler (its private RNG stream) cannot be
                     reproduced, so this file restates the *product's* sampling/solver contract (include/b200slam.h)
                     in numpy float64, and the tests compare three ways: CUDA vs this restatement (same seed: same
                     winning sample, same mask), CUDA vs cv2.findFundamentalMat (statistical: consensus size and
                     inlier-set overlap), and against the scene's ground-truth inliers.
* `reproject_and_match_2d3d` lives in oracle/pnp.py.
"""
from __future__ import annotations

import math

import numpy as np

M64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------------------------------
# cv2.remap, INTER_LINEAR, u8, CV_32FC1 maps, BORDER_CONSTANT(0)      (OpenCV imgwarp.cpp: remap / initInterTab2D)
# --------------------------------------------------------------------------------------------------------------
def remap_weight_table() -> np.ndarray:
    t = np.zeros((32, 32, 4), np.int16)
    for i in range(32):
        for j in range(32):
            vy = np.array([1.0 - i / 32.0, i / 32.0], np.float32)
            vx = np.array([1.0 - j / 32.0, j / 32.0], np.float32)
            f = (vy[:, None] * vx[None, :]).astype(np.float32).reshape(4)
            it = np.clip(np.rint(f * np.float32(32768)).astype(np.int64), -32768, 32767)
            s = int(it.sum())
            if s != 32768:
                d = 32768 - s
                it[int(it.argmax()) if d < 0 else int(it.argmin())] += d
            t[i, j] = it
    return t.reshape(1024, 4)


def remap_bgr_u8(src: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    H, W = src.shape[:2]
    sx = np.rint(mapx.astype(np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(mapy.astype(np.float32) * np.float32(32)).astype(np.int64)
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)
    w = remap_weight_table()[((sy & 31) << 5) | (sx & 31)].astype(np.int64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        return src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64) * ok[..., None]

    acc = (tap(iy, ix) * w[..., 0:1] + tap(iy, ix + 1) * w[..., 1:2]
           + tap(iy + 1, ix) * w[..., 2:3] + tap(iy + 1, ix + 1) * w[..., 3:4])
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------------------------
# 7-point RANSAC
# --------------------------------------------------------------------------------------------------------------
def _splitmix64(state: int):
    state = (state + 0x9E3779B97F4A7C15) & M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return state, z ^ (z >> 31)


def sample_indices(seed: int, h: int, n: int) -> list:
    """Seven distinct indices of hypothesis h (rejection on duplicates), include/b200slam.h `b2s_fm_ransac`."""
    st = (seed ^ ((0xD1B54A32D192ED03 * (h + 1)) & M64)) & M64
    idx = []
    while len(idx) < 7:
        st, z = _splitmix64(st)
        c = z % n
        if c not in idx:
            idx.append(int(c))
    return idx


def hartley(pts: np.ndarray):
    c = pts.astype(np.float64).mean(axis=0)
    d = np.sqrt(((pts.astype(np.float64) - c) ** 2).sum(axis=1)).mean()
    s = math.sqrt(2.0) / d if d > 1e-12 else 1.0
    return c, s


def _cubic_roots(c3, c2, c1, c0):
    """Real roots of the cubic (numpy companion-matrix roots, imaginary part below 1e-9 of the magnitude)."""
    co = np.array([c3, c2, c1, c0], np.float64)
    big = np.abs(co).max()
    if big == 0:
        return []
    while len(co) > 1 and abs(co[0]) <= 1e-12 * big:
        co = co[1:]
    if len(co) < 2:
        return []
    r = np.roots(co)
    return [float(z.real) for z in r if abs(z.imag) <= 1e-9 * max(1.0, abs(z))]


def seven_point(q: np.ndarray) -> list:
    """q [7,4] = (x1, y1, x2, y2) normalised.  All F (3x3) with x2^T F x1 = 0 and det F = 0."""
    x1, y1, x2, y2 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    A = np.stack([x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, np.ones(7)], axis=1)
    _, s, vt = np.linalg.svd(A)
    if s[6] <= 1e-10 * max(s[0], 1e-300):
        return []
    f1, f2 = vt[7].reshape(3, 3), vt[8].reshape(3, 3)
    G, D = f2, f1 - f2
    det = np.linalg.det
    c0, c3 = det(G), det(D)
    c1 = sum(det(np.where(np.arange(3)[:, None] == r, D, G)) for r in range(3))
    c2 = sum(det(np.where(np.arange(3)[:, None] == r, G, D)) for r in range(3))
    out = []
    for lam in _cubic_roots(c3, c2, c1, c0):
        F = G + lam * D
        if np.isfinite(F).all() and np.linalg.norm(F) > 0:
            out.append(F)
    return out


def fm_error(F: np.ndarray, p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    """OpenCV FMEstimatorCallback::computeError: max of the two squared point-to-epipolar-line distances."""
    p1, p2 = p1.astype(np.float64), p2.astype(np.float64)
    h1 = np.concatenate([p1, np.ones((len(p1), 1))], axis=1)
    h2 = np.concatenate([p2, np.ones((len(p2), 1))], axis=1)
    l2 = h1 @ F.T          # F x1
    l1 = h2 @ F            # F^T x2
    d2 = (h2 * l2).sum(axis=1)
    d1 = (h1 * l1).sum(axis=1)
    return np.maximum(d1 * d1 / (l1[:, 0] ** 2 + l1[:, 1] ** 2), d2 * d2 / (l2[:, 0] ** 2 + l2[:, 1] ** 2))


def denormalise(Fn: np.ndarray, c1, s1, c2, s2) -> np.ndarray:
    T1 = np.array([[s1, 0, -s1 * c1[0]], [0, s1, -s1 * c1[1]], [0, 0, 1.0]])
    T2 = np.array([[s2, 0, -s2 * c2[0]], [0, s2, -s2 * c2[1]], [0, 0, 1.0]])
    F = T2.T @ Fn @ T1
    return F / np.linalg.norm(F)


def hypothesis_models(pts1: np.ndarray, pts2: np.ndarray, seed: int, h: int) -> list:
    """Pixel-coordinate unit-norm candidate models of sample h (unordered: compare as sets up to sign)."""
    n = len(pts1)
    (c1, s1), (c2, s2) = hartley(pts1), hartley(pts2)
    idx = sample_indices(seed, h, n)
    q = np.concatenate([(pts1[idx].astype(np.float64) - c1) * s1, (pts2[idx].astype(np.float64) - c2) * s2], axis=1)
    return [denormalise(Fn, c1, s1, c2, s2) for Fn in seven_point(q)]


def fm_ransac(pts1: np.ndarray, pts2: np.ndarray, thresh: float, n_hyp: int, seed: int = 0):
    """Best-consensus model over n_hyp samples.  Returns (F or None, mask u8 [n], count, winning sample index)."""
    t2 = float(thresh) ** 2
    best = (-1, -1, None, None)
    for h in range(n_hyp):
        for F in hypothesis_models(pts1, pts2, seed, h):
            inl = fm_error(F, pts1, pts2) <= t2
            c = int(inl.sum())
            if c > best[0]:
                best = (c, h, F, inl)
    if best[0] < 7:
        return None, np.zeros(len(pts1), np.uint8), 0, -1
    return best[2], best[3].astype(np.uint8), best[0], best[1]


def two_view_scene(n: int, outlier_frac: float, noise_px: float, seed: int, W: int = 1241, H: int = 376):
    """Synthetic correspondences of a forward-moving KITTI-like camera: returns (pts1, pts2 f32 [n,2], gt_inlier bool)."""
    rng = np.random.default_rng(seed)
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.2], [0, 0, 1.0]])
    X = np.stack([rng.uniform(-15, 15, 4 * n), rng.uniform(-3, 3, 4 * n), rng.uniform(4, 60, 4 * n)], axis=1)
    ang = 0.03
    R = np.array([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]])
    t = np.array([0.15, -0.02, -0.9])
    a = (K @ X.T).T
    b = (K @ (X @ R.T + t).T).T
    p1, p2 = a[:, :2] / a[:, 2:3], b[:, :2] / b[:, 2:3]
    ok = (a[:, 2] > 0.5) & (b[:, 2] > 0.5) & (p1[:, 0] >= 0) & (p1[:, 0] < W) & (p1[:, 1] >= 0) & (p1[:, 1] < H) \
        & (p2[:, 0] >= 0) & (p2[:, 0] < W) & (p2[:, 1] >= 0) & (p2[:, 1] < H)
    p1, p2 = p1[ok][:n], p2[ok][:n]
    assert len(p1) == n, "scene too sparse"
    p1 = p1 + rng.normal(0, noise_px, p1.shape)
    p2 = p2 + rng.normal(0, noise_px, p2.shape)
    gt = np.ones(n, bool)
    bad = rng.choice(n, int(round(outlier_frac * n)), replace=False)
    p2[bad] = np.stack([rng.uniform(0, W, len(bad)), rng.uniform(0, H, len(bad))], axis=1)
    gt[bad] = False
    return p1.astype(np.float32), p2.astype(np.float32), gt
