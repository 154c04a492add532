"""CPU: pins of oracle/geometry.py against the reference's own implementation (OpenCV, which the reference calls at
slam/monocular/main_revamped.py:313-324 and slam/core/features_utils.py:193-194) and host-side checks of the library
boundary for the rows behind the matcher."""
import ctypes as C
import re

import cv2
import numpy as np

from oracle import geometry as G


def _kitti_maps(W=1241, H=376):
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.2], [0, 0, 1.0]])
    D = np.array([-0.28, 0.07, 0.0002, 0.0001, 0.0])
    nK, _ = cv2.getOptimalNewCameraMatrix(K, D, (W, H), alpha=0, newImgSize=(W, H))
    return cv2.initUndistortRectifyMap(K, D, None, nK, (W, H), cv2.CV_32FC1)


def test_remap_restatement_is_bit_exact_with_cv2():
    rng = np.random.default_rng(0)
    H, W = 376, 1241
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    mx, my = _kitti_maps(W, H)
    assert np.array_equal(cv2.remap(src, mx, my, cv2.INTER_LINEAR), G.remap_bgr_u8(src, mx, my))
    # arbitrary maps incl. out-of-image taps on every border and exact integer / half positions
    mx2 = rng.random((97, 131), dtype=np.float32) * (W + 40) - 20
    my2 = rng.random((97, 131), dtype=np.float32) * (H + 40) - 20
    mx2[0, :8] = [-1.0, -0.5, 0.0, 0.015625, W - 1, W - 0.5, W, 5.5]
    my2[0, :8] = [-1.0, -0.5, 0.0, 0.015625, H - 1, H - 0.5, H, 7.484375]
    assert np.array_equal(cv2.remap(src, mx2, my2, cv2.INTER_LINEAR), G.remap_bgr_u8(src, mx2, my2))


def test_remap_weight_table_sums():
    t = G.remap_weight_table().astype(np.int64)
    assert t.shape == (1024, 4) and (t.sum(axis=1) == 32768).all() and t[0].tolist() == [32767, 1, 0, 0]   # int16 saturation, as OpenCV


def test_seven_point_models_satisfy_their_sample():
    p1, p2, _ = G.two_view_scene(200, 0.0, 0.0, seed=3)
    for h in range(20):
        idx = G.sample_indices(11, h, len(p1))
        assert len(set(idx)) == 7 and all(0 <= i < len(p1) for i in idx)
        models = G.hypothesis_models(p1, p2, 11, h)
        assert 1 <= len(models) <= 3
        for F in models:
            assert abs(np.linalg.det(F)) < 1e-9 and abs(np.linalg.norm(F) - 1) < 1e-12
            assert G.fm_error(F, p1[idx], p2[idx]).max() < 1e-6      # the seven pairs lie on their epipolar lines


def test_fm_error_matches_opencv_definition():
    """computeError of OpenCV's FM estimator, checked through cv2 itself: every point cv2 flags as inlier has
    error <= thresh^2 under cv2's own F and every other point has a larger one."""
    p1, p2, _ = G.two_view_scene(600, 0.35, 0.2, seed=5)
    F, mask = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
    err = G.fm_error(F, p1, p2)
    m = mask.ravel().astype(bool)
    border = np.abs(err - 1.0) < 1e-3                  # cv2 evaluates in float32
    assert ((err <= 1.0) == m)[~border].all()


def test_ransac_restatement_against_cv2_and_ground_truth():
    for n, frac, noise, seed in [(800, 0.3, 0.15, 1), (400, 0.5, 0.2, 2)]:
        p1, p2, gt = G.two_view_scene(n, frac, noise, seed)
        F, mask, cnt, h = G.fm_ransac(p1, p2, 1.0, 192, seed=0)
        m = mask.astype(bool)
        _, mc = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
        mc = mc.ravel().astype(bool)
        assert cnt == m.sum() and cnt >= 0.95 * mc.sum()
        assert (m & gt).sum() >= 0.97 * m.sum()                        # precision against the true inliers
        assert (m & mc).sum() / (m | mc).sum() >= 0.85                 # overlap with cv2's consensus set
        assert h >= 0 and abs(np.linalg.det(F)) < 1e-9


def test_header_declares_what_the_library_exports():
    import b200slam._lib as L
    hdr = open(L._HERE + "/../include/b200slam.h").read()
    names = set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", hdr))
    for name in sorted(names):
        assert hasattr(L.lib, name), f"{name} declared in include/b200slam.h but not exported"
    for name in ("b2s_fm_create", "b2s_fm_ransac", "b2s_fm_ransac_host", "b2s_remap_create", "b2s_remap_bgr",
                 "b2s_remap_bgr_host", "b2s_aliked_set_undistort"):
        assert name in names and name in L.SIGNATURES


def test_geometry_handles_fail_loudly_without_a_device():
    import torch
    import b200slam._lib as L
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    assert L.lib.b2s_fm_create(0, 1024, 256, C.byref(h)) == -4            # B2S_ENODEV
    assert b"no CPU fallback" in L.lib.b2s_last_error()
    mx = np.zeros((4, 4), np.float32)
    assert L.lib.b2s_remap_create(0, mx.ctypes.data, mx.ctypes.data, 4, 4, 4, 4, C.byref(h)) == -4
    from b200slam import features_utils as fu
    kps = [cv2.KeyPoint(float(i), float(i % 7), 1) for i in range(20)]
    ms = [cv2.DMatch(i, i, 0.0) for i in range(20)]
    assert fu.filter_matches_ransac(kps, kps, ms[:5], 1.0) == ms[:5]      # < 8 matches: returned unchanged (reference :188-189)
    # default = the reference's own cv2 body: works without a GPU (ORB/SIFT-only users) ...
    p1, p2, _ = G.two_view_scene(200, 0.2, 0.2, 3)
    kpa = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p1]; kpb = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p2]
    msa = [cv2.DMatch(i, i, 0.0) for i in range(200)]
    out = fu.filter_matches_ransac(kpa, kpb, msa, 1.0)
    _, mask = cv2.findFundamentalMat(p1.astype(np.float32), p2.astype(np.float32), cv2.FM_RANSAC, 1.0, 0.99)
    assert [m.queryIdx for m in out] == np.flatnonzero(mask.ravel()).tolist() and 8 <= len(out) < 200
    # ... the GPU estimator is opt-in and has no CPU fallback
    try:
        fu.filter_matches_ransac_gpu(kps, kps, ms, 1.0)
        raise AssertionError("expected a RuntimeError without CUDA")
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)


def test_pnp_restatement_matches_reference_fixture():
    """oracle/pnp.py against tests/golden/pnp_reproj.npz = outputs of the reference's own reproject_and_match_2d3d
    (generated by tests/golden/make_golden_pnp.py from /root/reference/slam/core/pnp_utils.py)."""
    import os
    from oracle import pnp as O
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pnp_reproj.npz"))
    for c in range(4):
        n_points, n_kps, seed, radius, max_l2, cosine = g[f"c{c}_cfg"].tolist()
        wm, K, Tcw, kps, des = O.tracking_scene(int(n_points), int(n_kps), int(seed))
        r = O.reproject_and_match_2d3d(wm, K, Tcw, kps, des, 1241, 376, radius_px=radius, max_l2=max_l2, use_cosine=bool(cosine))
        assert len(r.mp_ids) >= 100
        assert r.mp_ids == g[f"c{c}_mp_ids"].tolist() and r.kp_indices == g[f"c{c}_kp_indices"].tolist()
        assert np.array_equal(r.pts3d, g[f"c{c}_pts3d"]) and np.array_equal(r.pts2d, g[f"c{c}_pts2d"])


def test_remap_restatement_random_shapes_against_cv2():
    """Randomised pin of the fixed-point remap restatement: odd sizes, destination != source size, maps that leave the
    image on every side, exact .5 / 1/32 sub-pixel positions (seeded, 40 cases)."""
    rng = np.random.default_rng(1234)
    for case in range(40):
        sh, sw = int(rng.integers(2, 70)), int(rng.integers(2, 90))
        dh, dw = int(rng.integers(1, 60)), int(rng.integers(1, 80))
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        kind = case % 4
        if kind == 0:      # arbitrary float positions, partly outside
            mx = (rng.random((dh, dw), dtype=np.float32) * (sw + 8) - 4).astype(np.float32)
            my = (rng.random((dh, dw), dtype=np.float32) * (sh + 8) - 4).astype(np.float32)
        elif kind == 1:    # positions on the 1/32 grid (exact weights) and half-way between grid points (rounding of cvRound)
            mx = (rng.integers(-64, 32 * sw + 64, (dh, dw)) / 32.0 + rng.choice([0.0, 1.0 / 64.0], (dh, dw))).astype(np.float32)
            my = (rng.integers(-64, 32 * sh + 64, (dh, dw)) / 32.0 + rng.choice([0.0, 1.0 / 64.0], (dh, dw))).astype(np.float32)
        elif kind == 2:    # affine warp
            ys, xs = np.mgrid[0:dh, 0:dw].astype(np.float32)
            a = rng.normal(0, 0.3, 4).astype(np.float32)
            mx = (1 + a[0]) * xs + a[1] * ys + np.float32(rng.normal(0, 3)); my = a[2] * xs + (1 + a[3]) * ys + np.float32(rng.normal(0, 3))
            mx, my = mx.astype(np.float32), my.astype(np.float32)
        else:              # far outside / huge coordinates
            mx = (rng.normal(0, 1e4, (dh, dw))).astype(np.float32); my = (rng.normal(0, 1e4, (dh, dw))).astype(np.float32)
        want = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
        got = G.remap_bgr_u8(src, mx, my)
        assert np.array_equal(want, got), f"case {case} kind {kind}: {(want != got).sum()} differing bytes"


def test_map_descriptor_packing_rules():
    """Host logic of the landmark mirror (pnp_utils.py:45-49, :112-123): last six observations, None entries skipped,
    no usable descriptor when the newest one is None."""
    from b200slam.pnp_utils import MapDescriptorMirror as M
    d = [np.full(128, i, np.float32) for i in range(9)]
    assert M._pack([]) == (0, None)
    assert M._pack([(0, 0, d[0]), (1, 1, None)])[0] == 0
    n, blk = M._pack([(i, i, d[i]) for i in range(9)])
    assert n == 6 and [int(r[0]) for r in blk] == [3, 4, 5, 6, 7, 8]
    obs = [(i, i, d[i]) for i in range(9)]; obs[5] = (5, 5, None)
    n, blk = M._pack(obs)
    assert n == 5 and [int(r[0]) for r in blk] == [3, 4, 6, 7, 8]
    try:
        M._pack([(0, 0, np.zeros(32, np.uint8))])
        raise AssertionError("binary descriptors must be rejected")
    except NotImplementedError:
        pass


def test_match_points_gather_equals_reference_list_comprehension():
    """`_match_points` == the reference's `np.float32([kp[m.queryIdx].pt for m in matches])` (features_utils.py:191-192)
    for cv2 lists and for the array-native containers."""
    from b200slam import features_utils as fu
    from b200slam.containers import DMatchArray, KeyPointArray
    rng = np.random.default_rng(0)
    p1 = (rng.random((50, 2)) * 500).astype(np.float32); p2 = (rng.random((60, 2)) * 500).astype(np.float32)
    pairs = np.stack([rng.integers(0, 50, 30), rng.integers(0, 60, 30)], axis=1).astype(np.int32)
    kp1 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p1]; kp2 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p2]
    ms = [cv2.DMatch(int(i), int(j), 0.0) for i, j in pairs]
    ref1 = np.float32([kp1[m.queryIdx].pt for m in ms]); ref2 = np.float32([kp2[m.trainIdx].pt for m in ms])
    a1, a2 = fu._match_points(kp1, kp2, ms)
    b1, b2 = fu._match_points(KeyPointArray(p1), KeyPointArray(p2), DMatchArray(pairs))
    assert np.array_equal(a1, ref1) and np.array_equal(a2, ref2) and np.array_equal(b1, ref1) and np.array_equal(b2, ref2)
    assert a1.dtype == np.float32 and a1.shape == (30, 2)
