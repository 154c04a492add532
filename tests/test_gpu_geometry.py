"""GPU: the rows behind the matcher through the C ABI - fundamental-matrix RANSAC against the numpy
restatement (same seed: same winning sample, same mask), against cv2.findFundamentalMat (the reference's call,
features_utils.py:193-194) and against ground truth; cv2.remap bit-exactness; ingest fused in front of the extractor."""
from types import SimpleNamespace

import cv2
import numpy as np
import pytest
import torch

from oracle import geometry as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ransac():
    from b200slam.geometry import FundamentalRansac
    return FundamentalRansac(max_points=4096, n_hyp=2048, seed=0)


def _same_model(Fa, Fb, tol=1e-7):
    Fa, Fb = Fa / np.linalg.norm(Fa), Fb / np.linalg.norm(Fb)
    return min(np.abs(Fa - Fb).max(), np.abs(Fa + Fb).max()) < tol


@pytest.mark.parametrize("n,frac,noise,seed", [(800, 0.3, 0.3, 1), (64, 0.2, 0.2, 3), (2048, 0.5, 0.5, 2)])
def test_ransac_matches_restatement(ransac, n, frac, noise, seed):
    """Same samples, same candidate models, same winner and mask as oracle/geometry.py (float64 both sides)."""
    from b200slam.geometry import FundamentalRansac
    p1, p2, _ = G.two_view_scene(n, frac, noise, seed)
    H = 96
    r = FundamentalRansac(max_points=4096, n_hyp=H, seed=7)
    F, mask = r.run_host(p1, p2, 1.0)
    Fo, mo, cnt_o, h_o = G.fm_ransac(p1, p2, 1.0, H, seed=7)
    # per-sample candidate sets (the solver): every oracle model appears among the device's and vice versa
    for h in (0, 1, 2, 17, H - 1):
        dm, dc = r.debug_models(h)
        om = G.hypothesis_models(p1, p2, 7, h)
        assert len(dm) == len(om)
        for Fm, c in zip(dm, dc):
            assert any(_same_model(Fm, Fq) for Fq in om)
            err = G.fm_error(Fm, p1, p2)
            assert abs(int((err <= 1.0).sum()) - int(c)) <= int((np.abs(err - 1.0) < 1e-9).sum())
    assert r.last_model_index // 3 == h_o and r.last_count == cnt_o
    assert _same_model(F, Fo)
    border = np.abs(G.fm_error(Fo, p1, p2) - 1.0) < 1e-9
    assert (mask.ravel() == mo)[~border].all() and int(mask.sum()) == r.last_count


@pytest.mark.parametrize("n,frac,noise,seed", [(800, 0.3, 0.15, 1), (1500, 0.5, 0.3, 2), (300, 0.6, 0.2, 5), (2048, 0.2, 0.1, 8)])
def test_ransac_against_cv2_and_ground_truth(ransac, n, frac, noise, seed):
    p1, p2, gt = G.two_view_scene(n, frac, noise, seed)
    F, mask = ransac.run_host(p1, p2, 1.0)
    m = mask.ravel().astype(bool)
    _, mc = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
    mc = mc.ravel().astype(bool)
    assert mask.shape == (n, 1) and mask.dtype == np.uint8 and F.shape == (3, 3)
    assert m.sum() >= 0.98 * mc.sum(), "consensus smaller than cv2's"
    assert (m & gt).sum() >= 0.97 * m.sum(), "false inliers"
    assert (m & gt).sum() >= (mc & gt).sum() - 0.02 * gt.sum(), "recall below cv2's"
    # properties of the returned model: rank 2, every flagged pair within the threshold, nothing else
    err = G.fm_error(F, p1, p2)
    assert abs(np.linalg.det(F / np.linalg.norm(F))) < 1e-9
    assert (err[m] <= 1.0 + 1e-9).all() and (err[~m] > 1.0 - 1e-9).all()


def test_ransac_low_noise_overlap_with_cv2(ransac):
    p1, p2, gt = G.two_view_scene(1200, 0.3, 0.05, 4)
    _, mask = ransac.run_host(p1, p2, 1.0)
    _, mc = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
    m, mc = mask.ravel().astype(bool), mc.ravel().astype(bool)
    assert (m & mc).sum() / (m | mc).sum() >= 0.97
    assert (m == gt).mean() >= 0.99


def test_ransac_device_resident_pairs(ransac):
    """pairs_dev indirection (the matcher's device `matches` feed RANSAC without a host round trip)."""
    p1, p2, _ = G.two_view_scene(700, 0.3, 0.2, 9)
    rng = np.random.default_rng(0)
    perm1, perm2 = rng.permutation(900), rng.permutation(900)
    k0 = np.zeros((900, 2), np.float32); k1 = np.zeros((900, 2), np.float32)
    k0[perm1[:700]] = p1; k1[perm2[:700]] = p2
    pairs = np.stack([perm1[:700], perm2[:700]], axis=1).astype(np.int32)
    mask_d, F_d, res_d = ransac.run_device(torch.from_numpy(k0).cuda(), torch.from_numpy(k1).cuda(),
                                           torch.from_numpy(pairs).cuda(), 700, 1.0)
    torch.cuda.synchronize()
    F, mask = ransac.run_host(p1, p2, 1.0)
    assert np.array_equal(mask_d.cpu().numpy(), mask.ravel()) and int(res_d[0]) == int(mask.sum())
    assert np.allclose(F_d.cpu().numpy().reshape(3, 3), F, atol=0, rtol=0)


def test_ransac_edge_cases(ransac):
    from b200slam import features_utils as fu
    kps = [cv2.KeyPoint(float(3 * i), float(i % 11), 1) for i in range(30)]
    ms = [cv2.DMatch(i, i, 0.0) for i in range(30)]
    assert fu.filter_matches_ransac_gpu(kps, kps, ms[:7], 1.0) == ms[:7]          # < 8: unchanged (reference :188-189)
    assert ransac.run_host(np.zeros((5, 2), np.float32), np.zeros((5, 2), np.float32)) == (None, None)
    # every correspondence identical -> every sample degenerate -> no model -> reference returns [] (mask is None)
    same = [cv2.KeyPoint(5.0, 5.0, 1)] * 30
    assert fu.filter_matches_ransac_gpu(same, same, ms, 1.0) == []
    # pure translation of a planar grid is consistent with many F: whatever is returned must satisfy its own mask
    out = fu.filter_matches_ransac_gpu(kps, kps, ms, 1.0)
    assert isinstance(out, list) and all(isinstance(m, cv2.DMatch) for m in out)
    # growth beyond the handle's capacity
    p1, p2, _ = G.two_view_scene(5000, 0.1, 0.1, 12)
    F, mask = ransac.run_host(p1, p2, 1.0)
    assert mask.sum() >= 0.85 * 4500 and ransac.max_points >= 5000


def test_filter_matches_ransac_default_is_reference_identical():
    """a11: the default `filter_matches_ransac` is the reference's body (features_utils.py:185-200): on the same inputs the
    surviving matches are IDENTICAL to `cv2.findFundamentalMat(pts1, pts2, FM_RANSAC, thresh, 0.99)`'s mask, scene by scene."""
    from b200slam import features_utils as fu
    from b200slam.containers import DMatchArray, KeyPointArray
    for seed, (n, out_frac, noise, thresh) in enumerate([(600, 0.3, 0.15, 1.0), (1300, 0.2, 0.3, 2.5), (2048, 0.1, 0.1, 1.0), (90, 0.4, 0.5, 3.0)]):
        p1, p2, _ = G.two_view_scene(n, out_frac, noise, 40 + seed)
        kp1 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p1]
        kp2 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p2]
        ms = [cv2.DMatch(i, i, 0.0) for i in range(n)]
        _, mask = cv2.findFundamentalMat(np.float32([kp1[m.queryIdx].pt for m in ms]), np.float32([kp2[m.trainIdx].pt for m in ms]),
                                         cv2.FM_RANSAC, thresh, 0.99)
        want = [m.queryIdx for m, ok in zip(ms, mask.ravel().astype(bool)) if ok]
        got = fu.filter_matches_ransac(kp1, kp2, ms, thresh)
        assert [m.queryIdx for m in got] == want and 8 <= len(want) < n
        got_a = fu.filter_matches_ransac(KeyPointArray(p1), KeyPointArray(p2), DMatchArray(np.stack([np.arange(n)] * 2, 1)), thresh)
        assert isinstance(got_a, DMatchArray) and got_a.queryIdx.tolist() == want


def test_filter_matches_ransac_gpu_opt_in(ransac):
    """The GPU estimator (opt-in): same signature / return type; consensus statistically equal to cv2's."""
    from b200slam import features_utils as fu
    from b200slam.containers import DMatchArray, KeyPointArray
    p1, p2, gt = G.two_view_scene(600, 0.3, 0.15, 21)
    kp1 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p1]
    kp2 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p2]
    ms = [cv2.DMatch(i, i, 0.0) for i in range(600)]
    fu.set_gpu_ransac(True)
    try:
        out = fu.filter_matches_ransac(kp1, kp2, ms, 1.0)
    finally:
        fu.set_gpu_ransac(False)
    ref = fu.filter_matches_ransac(kp1, kp2, ms, 1.0)
    assert isinstance(out, list) and isinstance(out[0], cv2.DMatch)
    so, sr = {m.queryIdx for m in out}, {m.queryIdx for m in ref}
    assert len(so) >= 0.98 * len(sr) and len(so & sr) / len(so | sr) >= 0.9
    assert np.mean([gt[i] for i in so]) >= 0.97
    out_a = fu.filter_matches_ransac_gpu(KeyPointArray(p1), KeyPointArray(p2), DMatchArray(np.stack([np.arange(600)] * 2, 1)), 1.0)
    assert isinstance(out_a, DMatchArray) and set(out_a.queryIdx.tolist()) == so


def _kitti_maps(W=1241, H=376):
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.2], [0, 0, 1.0]])
    D = np.array([-0.28, 0.07, 0.0002, 0.0001, 0.0])
    nK, _ = cv2.getOptimalNewCameraMatrix(K, D, (W, H), alpha=0, newImgSize=(W, H))
    return cv2.initUndistortRectifyMap(K, D, None, nK, (W, H), cv2.CV_32FC1)


def test_remap_bit_exact_with_cv2():
    from b200slam import synth
    from b200slam.geometry import FrameUndistorter
    img = synth.frame(3, 376, 1241)
    mx, my = _kitti_maps()
    und = FrameUndistorter(mx, my)
    out = und.remap(img)
    assert np.array_equal(out, cv2.remap(img, mx, my, cv2.INTER_LINEAR))
    assert np.array_equal(out, G.remap_bgr_u8(img, mx, my))
    assert np.array_equal(und.remap_to_device(img).cpu().numpy(), out)
    # arbitrary maps with out-of-image taps, destination size != source size
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 256, (200, 333, 3), dtype=np.uint8)
    mx2 = rng.random((97, 131), dtype=np.float32) * (333 + 40) - 20
    my2 = rng.random((97, 131), dtype=np.float32) * (200 + 40) - 20
    mx2[0, :8] = [-1.0, -0.5, 0.0, 0.015625, 332, 332.5, 333, 5.5]
    my2[0, :8] = [-1.0, -0.5, 0.0, 0.015625, 199, 199.5, 200, 7.484375]
    und2 = FrameUndistorter(mx2, my2, src_shape=(200, 333))
    assert np.array_equal(und2.remap(noise), cv2.remap(noise, mx2, my2, cv2.INTER_LINEAR))
    with pytest.raises(ValueError):
        und2.remap(img)


def test_ingest_fused_in_front_of_the_extractor(aliked_state):
    """feature_extractor(raw frame) with an attached undistorter == feature_extractor(cv2.remap(raw frame))."""
    from b200slam import features_utils as fu, frontend, synth
    from b200slam.geometry import FrameUndistorter
    args = SimpleNamespace(use_lightglue=True, max_features=1024)
    det = frontend.ALIKED(max_num_keypoints=1024, weights=aliked_state)
    img = synth.frame(5, 376, 1241)
    mx, my = _kitti_maps()
    kp_ref, de_ref = fu.feature_extractor(args, cv2.remap(img, mx, my, cv2.INTER_LINEAR), det)
    det.set_undistort(FrameUndistorter(mx, my))
    kp, de = fu.feature_extractor(args, img, det)
    det.set_undistort(None)
    assert [k.pt for k in kp] == [k.pt for k in kp_ref] and np.array_equal(de, de_ref)
    kp2, _ = fu.feature_extractor(args, img, det)                  # detached again: the raw frame gives other keypoints
    assert [k.pt for k in kp2] != [k.pt for k in kp_ref]


# ---- reproject_and_match_2d3d (pnp_utils.py:224-304) -------------------------------------------------------------
def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "pnp_reproj.npz"))


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_reproject_and_match_against_reference_fixture(case):
    """Identical landmark ids / keypoint indices / coordinates as the reference's own function (committed fixture)
    and as the numpy restatement on the same seeded scene."""
    from b200slam import pnp_utils as P
    from oracle import pnp as O
    g = _golden()
    n_points, n_kps, seed, radius, max_l2, cosine = g[f"c{case}_cfg"].tolist()
    wm, K, Tcw, kps, des = O.tracking_scene(int(n_points), int(n_kps), int(seed))
    r = P.reproject_and_match_2d3d(wm, K, Tcw, kps, des, 1241, 376, radius_px=radius, max_l2=max_l2, use_cosine=bool(cosine))
    assert isinstance(r, P.Matches2D3D) and r.pts3d.dtype == np.float32 and r.pts2d.dtype == np.float32
    assert r.mp_ids == g[f"c{case}_mp_ids"].tolist()
    assert r.kp_indices == g[f"c{case}_kp_indices"].tolist()
    assert np.array_equal(r.pts3d, g[f"c{case}_pts3d"]) and np.array_equal(r.pts2d, g[f"c{case}_pts2d"])


def test_reproject_and_match_map_updates_and_edge_cases():
    import cv2 as _cv2
    from b200slam import pnp_utils as P
    from oracle import pnp as O
    wm, K, Tcw, kps, des = O.tracking_scene(800, 1024, 11)
    m = P.ReprojectionMatcher(max_points=256, max_kps=256)         # both capacities must grow
    a = m.match(wm, K, Tcw, kps, des, 1241, 376, 12.0, 0.8)
    o = O.reproject_and_match_2d3d(wm, K, Tcw, kps, des, 1241, 376, 12.0, 0.8)
    assert a.mp_ids == o.mp_ids and a.kp_indices == o.kp_indices and len(a.mp_ids) > 100
    # the map changes between frames: new observations, culled and new landmarks, moved positions (after BA)
    rng = np.random.default_rng(0)
    ids = list(wm.points)
    for pid in ids[::7]:
        del wm.points[pid]
    for pid in ids[1::5]:
        if pid in wm.points:
            d = rng.normal(size=128).astype(np.float32); d /= np.linalg.norm(d)
            wm.points[pid].observations.append((99, 0, d))
    for pid in ids[2::9]:
        if pid in wm.points:
            wm.points[pid].position = wm.points[pid].position + rng.normal(0, 0.05, 3)
    for k in range(40):
        src = wm.points[ids[3 + 7 * k + 1]] if ids[3 + 7 * k + 1] in wm.points else None
        if src is not None:
            wm.points[900000 + k] = O._MP(900000 + k, src.position + 0.01, list(src.observations))
    b = m.match(wm, K, Tcw, kps, des, 1241, 376, 12.0, 0.8)
    o2 = O.reproject_and_match_2d3d(wm, K, Tcw, kps, des, 1241, 376, 12.0, 0.8)
    assert b.mp_ids == o2.mp_ids and b.kp_indices == o2.kp_indices and b.mp_ids != a.mp_ids
    # cv2.KeyPoint list input, as the tracking loop passes it (main_revamped.py:460)
    kp_list = [_cv2.KeyPoint(float(x), float(y), 1) for x, y in kps]
    c = P.reproject_and_match_2d3d(wm, K, Tcw, kp_list, des, 1241, 376, radius_px=12.0)
    assert c.mp_ids == o2.mp_ids and c.kp_indices == o2.kp_indices
    # empties (pnp_utils.py:239-247)
    e = P.reproject_and_match_2d3d(wm, K, Tcw, kp_list, des[:0], 1241, 376)
    assert e.pts3d.shape == (0, 3) and e.pts2d.shape == (0, 2) and e.kp_indices == [] and e.mp_ids == []
    assert P.reproject_and_match_2d3d(O._Map(), K, Tcw, kp_list, des, 1241, 376).mp_ids == []
    assert P.reproject_and_match_2d3d(wm, K, Tcw, [], des, 1241, 376).mp_ids == []
    # camera looking away: nothing projects into the image
    Tb = Tcw.copy(); Tb[:3, :3] = Tcw[:3, :3] @ np.diag([-1.0, 1.0, -1.0]); Tb[:3, 3] = [0, 0, -200.0]
    assert P.reproject_and_match_2d3d(wm, K, Tb, kps, des, 1241, 376).mp_ids == O.reproject_and_match_2d3d(wm, K, Tb, kps, des, 1241, 376).mp_ids
    # a dense cluster: 200 keypoints inside one landmark's search window (more than the 64 candidate slots).  Only
    # keypoints that pass the descriptor gate are stored, so the reference's result is reproduced - no error, no limit
    # on --proj_radius (the reference has none, pnp_utils.py:268-295)
    wm2, K2, T2, kps2, des2 = O.tracking_scene(300, 512, 5)
    usable = [p for p in wm2.points.values() if p.observations and p.observations[-1][2] is not None]
    uv, _ = O.project_points(K2, T2, np.asarray([p.position for p in usable]))
    vis = uv[(uv[:, 0] > 50) & (uv[:, 0] < 1100) & (uv[:, 1] > 50) & (uv[:, 1] < 300)][0]
    kps2[:200] = vis + rng.normal(0, 0.5, (200, 2)).astype(np.float32)
    for radius in (12.0, 40.0):
        c2 = P.reproject_and_match_2d3d(wm2, K2, T2, kps2, des2, 1241, 376, radius_px=radius)
        o3 = O.reproject_and_match_2d3d(wm2, K2, T2, kps2, des2, 1241, 376, radius_px=radius)
        assert c2.mp_ids == o3.mp_ids and c2.kp_indices == o3.kp_indices and len(o3.mp_ids) > 50
    # ... and when more than 64 keypoints of a window DO pass the gate (identical descriptors) and a landmark loses all of
    # them to earlier landmarks, the matcher retries with the widest candidate list instead of failing
    des3 = des2.copy(); des3[:200] = des3[0]
    c3 = P.reproject_and_match_2d3d(wm2, K2, T2, kps2, des3, 1241, 376, radius_px=12.0, max_l2=2.5)
    o4 = O.reproject_and_match_2d3d(wm2, K2, T2, kps2, des3, 1241, 376, radius_px=12.0, max_l2=2.5)
    assert c3.mp_ids == o4.mp_ids and c3.kp_indices == o4.kp_indices


# ---- size-independent properties at the BASELINE sizes (2048 keypoints, 1241x376) ---------------------------------
def test_ransac_properties_full_size(ransac):
    """Permutation equivariance is not expected of a sampled estimator, but these are: determinism for a fixed seed,
    invariance of the consensus under a common translation of both images' coordinates (Hartley normalisation), and
    the mask always being the exact consensus set of the returned F."""
    p1, p2, gt = G.two_view_scene(2048, 0.35, 0.2, 31)
    F1, m1 = ransac.run_host(p1, p2, 1.0, seed=5)
    F2, m2 = ransac.run_host(p1, p2, 1.0, seed=5)
    assert np.array_equal(m1, m2) and np.array_equal(F1, F2)
    F3, m3 = ransac.run_host(p1, p2, 1.0, seed=6)                      # another sample stream: a different, equally good model
    assert abs(int(m3.sum()) - int(m1.sum())) <= 0.03 * m1.sum()
    shift = np.float32([64.0, -32.0])
    _, ms = ransac.run_host(p1 + shift, p2 + shift, 1.0, seed=5)       # exactly representable shift: same samples, same geometry
    assert (ms == m1).mean() >= 0.95 and abs(int(ms.sum()) - int(m1.sum())) <= 0.03 * m1.sum()
    err = G.fm_error(F1, p1, p2)
    assert (err[m1.ravel() == 1] <= 1.0 + 1e-9).all() and (err[m1.ravel() == 0] > 1.0 - 1e-9).all()
    # a looser threshold can only grow the consensus of the winning model's own mask
    F4, m4 = ransac.run_host(p1, p2, 2.0, seed=5)
    assert m4.sum() >= m1.sum()


def test_remap_identity_and_shift_properties():
    from b200slam import synth
    from b200slam.geometry import FrameUndistorter
    img = synth.frame(7, 376, 1241)
    H, W = img.shape[:2]
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    assert np.array_equal(FrameUndistorter(xs, ys).remap(img), img)                       # identity maps reproduce the frame
    out = FrameUndistorter(xs + 3.0, ys - 2.0).remap(img)                                 # integer shift = crop + zero border
    assert np.array_equal(out[2:, :W - 3], img[:H - 2, 3:]) and not out[:2].any() and not out[:, W - 3:].any()
    half = FrameUndistorter(xs + 0.5, ys).remap(img)                                      # half-pixel: rounded mean of neighbours
    ref = ((img[:, :-1].astype(np.int32) + img[:, 1:].astype(np.int32) + 1) >> 1).astype(np.uint8)
    assert np.array_equal(half[:, :W - 1], ref)


def test_reproject_and_match_permutation_property():
    """Relabelling the current frame's keypoints permutes the returned keypoint indices and nothing else."""
    from b200slam import pnp_utils as P
    from oracle import pnp as O
    wm, K, Tcw, kps, des = O.tracking_scene(3000, 2048, 41)
    a = P.reproject_and_match_2d3d(wm, K, Tcw, kps, des, 1241, 376, radius_px=12.0)
    perm = np.random.default_rng(3).permutation(len(kps))
    b = P.reproject_and_match_2d3d(wm, K, Tcw, kps[perm], des[perm], 1241, 376, radius_px=12.0)
    assert len(a.mp_ids) > 500 and a.mp_ids == b.mp_ids
    assert [int(perm[i]) for i in b.kp_indices] == a.kp_indices and np.array_equal(a.pts2d, b.pts2d) and np.array_equal(a.pts3d, b.pts3d)
    # every assignment honours the reference's gates: inside the window, within the descriptor threshold, keypoints used once
    uv, z = O.project_points(K, Tcw, a.pts3d.astype(np.float64))
    assert (np.linalg.norm(uv - a.pts2d, axis=1) <= 12.0 + 1e-3).all() and len(set(a.kp_indices)) == len(a.kp_indices)


# ---------------------------------------------------------------------------------------------------------------
# f1 / a11: OpenCV's RANSAC reproduced on the GPU (b2s_fm_cv_ransac_host) - identical to cv2, not statistically close
# ---------------------------------------------------------------------------------------------------------------
def _cv_scenes():
    sizes = [20, 60, 150, 300, 700, 1300, 2048, 15, 33, 500, 4000, 97]
    for seed in range(36):
        n = sizes[seed % len(sizes)]
        yield seed, n, [0.1, 0.3, 0.5, 0.2, 0.6, 0.8][seed % 6], [0.2, 0.5, 1.0][seed % 3], [1.0, 1.0, 3.0, 0.5, 2.5][seed % 5]


def test_fm_cv_ransac_is_identical_to_cv2(ransac):
    """Mask equality and F to 1e-8 against `cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, 0.99)` - the reference's
    exact call (features_utils.py:195-196) - on 36 seeded scenes (15 ... 4000 matches, 10 ... 80 % outliers)."""
    for seed, n, of, noise, thresh in _cv_scenes():
        p1, p2, _ = G.two_view_scene(n, of, noise, seed=500 + seed)
        Fc, mc = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, thresh, 0.99)
        Fg, mg = ransac.run_cv(p1, p2, thresh, 0.99)
        assert mg is not None and np.array_equal(mc, mg), (seed, n, int(mc.sum()), None if mg is None else int(mg.sum()), ransac.last_info)
        assert np.abs(Fc - Fg).max() <= 1e-8 * np.abs(Fc).max(), (seed, n)
        assert ransac.last_count == int(mc.sum())


def test_fm_cv_ransac_golden_fixtures(ransac):
    """The committed masks / F were produced by cv2 (tests/golden/make_golden_fm_cv.py); incl. an integer grid with
    duplicated points (getSubset redraws, checkSubset rejections) and 4000 correspondences."""
    from helpers import fm_cv_golden_cases
    for c, p1, p2, thresh, F, mask in fm_cv_golden_cases():
        Fg, mg = ransac.run_cv(p1, p2, thresh, 0.99)
        assert np.array_equal(mask, mg), c
        assert np.abs(F - Fg).max() <= 1e-8 * np.abs(F).max(), c


def test_fm_cv_ransac_matches_restatement_bookkeeping(ransac):
    """Winning iteration / model / final iteration bound equal the restatement's (oracle/cv_ransac.py)."""
    from oracle import cv_ransac as R
    for seed in (1, 2, 3):
        p1, p2, _ = G.two_view_scene(400, 0.5, 0.5, seed=700 + seed)
        _, mg = ransac.run_cv(p1, p2, 1.0, 0.99)
        _, mo, info = R.ransac_run(*R._prep(p1, p2), 1.0)
        assert np.array_equal(mg.ravel(), mo) and ransac.last_info[:3] == (info["winner"][0], info["winner"][1], info["niters"])


def test_filter_matches_ransac_gpu_cv2_mode_is_reference_identical():
    """`filter_matches_ransac` in "gpu_cv2" mode keeps exactly the matches the reference keeps (lists and array-native
    containers); fewer than 15 matches are OpenCV's LMedS branch and go to cv2; fewer than 8 pass through."""
    from b200slam import features_utils as fu
    from b200slam.containers import DMatchArray, KeyPointArray
    fu.set_ransac_mode("gpu_cv2")
    try:
        for seed, (n, out_frac, noise, thresh) in enumerate([(600, 0.3, 0.15, 1.0), (1300, 0.2, 0.3, 2.5), (2048, 0.1, 0.1, 1.0),
                                                             (90, 0.4, 0.5, 3.0), (12, 0.2, 0.3, 1.0), (7, 0.0, 0.1, 1.0)]):
            p1, p2, _ = G.two_view_scene(n, out_frac, noise, 40 + seed)
            kp1 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p1]
            kp2 = [cv2.KeyPoint(float(x), float(y), 1) for x, y in p2]
            ms = [cv2.DMatch(i, i, 0.0) for i in range(n)]
            want = [m.queryIdx for m in fu.filter_matches_ransac_cv2(kp1, kp2, ms, thresh)]
            got = fu.filter_matches_ransac(kp1, kp2, ms, thresh)
            assert [m.queryIdx for m in got] == want
            got_a = fu.filter_matches_ransac(KeyPointArray(p1), KeyPointArray(p2), DMatchArray(np.stack([np.arange(n)] * 2, 1)), thresh)
            assert isinstance(got_a, DMatchArray) and got_a.queryIdx.tolist() == want
        # degenerate input (no valid 7-point sample): cv2 returns F = None with an UNINITIALISED mask, which the reference then
        # reads; the GPU path returns no model and filter_matches_ransac keeps nothing (features_utils.py:197-198's branch)
        same = [cv2.KeyPoint(10.0, 20.0, 1)] * 30
        assert fu.filter_matches_ransac(same, same, [cv2.DMatch(i, i, 0.0) for i in range(30)], 1.0) == []
    finally:
        fu.set_ransac_mode("cv2")
