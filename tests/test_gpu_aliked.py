"""GPU parity: ALIKED extractor (C-ABI b2s_aliked_extract[_host]) vs the CPU oracle.
Bar: identical keypoint index sets; dense maps / descriptors within 1e-3 relative (measured ~1e-6);
NMS bit-exact on identical input."""
import numpy as np
import pytest
import torch

import oracle
from b200slam import weights, synth
from helpers import rel_err, aliked_decision_margins

pytestmark = pytest.mark.gpu


def _pair(model="aliked-n16", max_kp=2048):
    from b200slam import frontend
    sd = weights.synthetic_aliked_state(model)
    ora = oracle.ALIKED(model_name=model, max_num_keypoints=max_kp).eval()
    ora.load_state_dict(sd, strict=True)
    ora.record_taps = True
    return ora, frontend.ALIKED(model_name=model, max_num_keypoints=max_kp, weights=sd)


def _pix_index(det, n, Hr, Wr):
    kn = det.debug("kp_norm").reshape(-1, 2)[:n]
    # integer NMS location = round(refined - residual) is not recoverable exactly; use floor(x+0.5) of the
    # refined position, which stays inside the 5x5 soft-argmax support centre +-0.5 for peaked maxima
    return kn


CASES = {"kitti_2048": (376, 1241, 2048, "aliked-n16"), "vga_upscale_1024": (480, 640, 1024, "aliked-n16"),
         "small_512": (240, 320, 512, "aliked-n16"), "n32_1080p_4096": (1080, 1920, 4096, "aliked-n32")}


@pytest.mark.parametrize("name", list(CASES))
def test_extract_parity(name):
    H, W, max_kp, model = CASES[name]
    ora, det = _pair(model, max_kp)
    img = synth.frame(3, H, W)
    fo = ora.extract(oracle.bgr_to_tensor(img))
    kp, de, sc = det.extract_host(img)                       # C-ABI, host buffers
    Hr, Wr, Hp, Wp = [int(v) for v in det.debug("geometry")]
    T = ora.taps
    assert (Hr, Wr) == tuple(T["score_map"].shape[-2:])
    # stage taps (tolerance 1e-3 relative to the tensor's max; observed ~1e-6)
    assert rel_err(det.debug("resized").reshape(3, Hr, Wr), T["resized"][0].numpy()) < 1e-5
    assert rel_err(det.debug("x1").reshape(16, Hp, Wp), T["x1"][0].numpy()) < 1e-4
    assert rel_err(det.debug("x2").reshape(32, Hp // 2, Wp // 2), T["x2"][0].numpy()) < 1e-4
    assert rel_err(det.debug("x3").reshape(Hp // 8, Wp // 8, 64).transpose(2, 0, 1), T["x3"][0].numpy()) < 1e-4
    assert rel_err(det.debug("x4").reshape(Hp // 32, Wp // 32, 128).transpose(2, 0, 1), T["x4"][0].numpy()) < 1e-4
    assert rel_err(det.debug("score_map").reshape(Hr, Wr), T["score_map"][0, 0].numpy()) < 1e-4
    assert rel_err(det.debug("feature_map").reshape(Hr, Wr, 128).transpose(2, 0, 1), T["feature_map"][0].numpy()) < 1e-4
    # NMS kernel is bit-exact against the oracle NMS of the same (GPU) score map
    sg = torch.from_numpy(det.debug("score_map").reshape(1, 1, Hr, Wr).copy())
    nms = oracle.simple_nms(sg, 2)[0, 0].numpy().copy()
    nms[:2] = 0; nms[-2:] = 0; nms[:, :2] = 0; nms[:, -2:] = 0
    assert np.array_equal(nms, det.debug("nms").reshape(Hr, Wr))
    # identical keypoint sets (compare positions; order may swap between near-tied scores)
    ko = fo["keypoints"][0].numpy()
    assert len(kp) == len(ko)
    key = lambda a: np.rint(a * 16).astype(np.int64)   # 1/16 px grid (positions agree to ~1e-4 px)
    so = {tuple(r): i for i, r in enumerate(key(ko).tolist())}
    idx = np.array([so.get(tuple(r), -1) for r in key(kp).tolist()])
    assert (idx >= 0).all(), f"{(idx < 0).sum()} GPU keypoints not in the oracle set"
    assert len(set(idx.tolist())) == len(ko)
    assert np.abs(kp - ko[idx]).max() < 1e-2
    # descriptors / scores of the same keypoints: 1e-3 relative
    assert rel_err(de, fo["descriptors"][0].numpy()[idx]) < 1e-3
    assert np.abs(np.linalg.norm(de, axis=1) - 1).max() < 1e-5
    assert rel_err(sc, fo["keypoint_scores"][0].numpy()[idx]) < 1e-3
    assert (idx == np.arange(len(idx))).mean() > 0.98          # order: score-descending like upstream


def test_float_chw_entry_equals_u8_entry_and_contract():
    """detector.extract(tensor) (features_utils.py:94) vs the fused u8 BGR entry; dict contract."""
    ora, det = _pair(max_kp=600)
    img = synth.frame(5, 200, 300)
    f_u8 = det.extract_bgr(img)
    f_f32 = det.extract(oracle.bgr_to_tensor(img))            # CPU tensor in, moved by the shim
    assert set(f_f32) == {"keypoints", "descriptors", "keypoint_scores", "image_size"}
    assert f_f32["keypoints"].shape == f_u8["keypoints"].shape and f_f32["keypoints"].shape[0] == 1
    assert torch.equal(f_f32["keypoints"], f_u8["keypoints"]) and torch.equal(f_f32["descriptors"], f_u8["descriptors"])
    assert f_f32["image_size"].tolist() == [[300.0, 200.0]]
    assert det.eval() is det and det.to("cuda") is det and next(det.parameters()).is_cuda


def test_untruncated_raster_order():
    """fewer candidates than n_limit -> upstream keeps raster (nonzero) order.  The input is a 6x up-sampled
    120x160 image, i.e. smooth score plateaus where a maximum / the 0.2 threshold can be decided by the last
    bit of the score.  Every keypoint found by only one side must be such a flip: its DECISION MARGIN (SURVEY.md 7,
    helpers.aliked_decision_margins) is below 1e-6; the common ones come in the same order."""
    ora, det = _pair(max_kp=-1)
    img = synth.frame(7, 120, 160)
    fo = ora.extract(oracle.bgr_to_tensor(img))
    kp, de, _ = det.extract_host(img)
    ko = fo["keypoints"][0].numpy()
    Hr, Wr = [int(v) for v in det.debug("geometry")][:2]
    mg = aliked_decision_margins(det.debug("score_map").reshape(Hr, Wr), ora.taps["score_map"][0, 0].numpy(), -1)
    print(f"decision margins: {mg['n_diff']} differing keypoints of {len(ko)}, max margin {mg['max_margin']:.3e}, "
          f"min |score-0.2| over maxima {mg['min_thr_margin']:.3e}")
    assert mg["max_margin"] < 1e-6, f"a keypoint differs with a real margin: {mg['margins']}"
    assert len(kp) < 20000 and abs(len(kp) - len(ko)) <= mg["n_diff"]
    key = lambda a: [tuple(r) for r in np.rint(a * 16).astype(np.int64).tolist()]   # noqa: E731
    so = {k: i for i, k in enumerate(key(ko))}
    idx = np.array([so.get(k, -1) for k in key(kp)])
    common = idx >= 0
    assert common.sum() >= len(ko) - mg["n_diff"]
    assert (np.diff(idx[common]) > 0).all(), "raster order must be preserved when nothing is truncated"
    assert np.abs(kp[common] - ko[idx[common]]).max() < 1e-2
    assert rel_err(de[common], fo["descriptors"][0].numpy()[idx[common]]) < 1e-3


def test_extract_batch_equals_per_frame_extraction():
    """b2s_aliked_extract_batch (frames spread over concurrent extractor lanes, forked from / joined into the caller's
    stream) gives bit for bit what per-frame b2s_aliked_extract calls give."""
    from b200slam import _lib
    _, det = _pair(max_kp=700)
    Hh, Ww = 240, 320
    imgs = [torch.from_numpy(synth.frame(40 + t, Hh, Ww)).cuda() for t in range(7)]
    ref = []
    for im in imgs:
        kp, de, sc, n = det.extract_device(im, _lib.IMG_BGR_U8_HWC, Hh, Ww, 3 * Ww)
        torch.cuda.synchronize()
        k = int(n)
        ref.append((kp[:k].clone(), de[:k].clone(), sc[:k].clone()))
    for lanes in (1, 3, 4):
        kp, de, sc, n = det.extract_batch_device(imgs, _lib.IMG_BGR_U8_HWC, Hh, Ww, 3 * Ww, lanes=lanes)
        torch.cuda.synchronize()
        for i, (rk, rd, rs) in enumerate(ref):
            k = int(n[i])
            assert k == len(rk) and torch.equal(kp[i, :k], rk) and torch.equal(de[i, :k], rd) and torch.equal(sc[i, :k], rs), (lanes, i)


def test_degenerate_images_do_not_crash():
    _, det = _pair(max_kp=256)
    for img in (np.zeros((64, 48, 3), np.uint8), np.full((100, 333, 3), 255, np.uint8)):
        kp, de, sc = det.extract_host(img)
        assert len(kp) <= 256 and np.isfinite(kp).all() and np.isfinite(de).all()
    with pytest.raises(Exception):
        det.extract_host(np.zeros((4, 4, 3), np.uint8))     # too small: error code, not a crash


def test_fp16x2_convolutions_agree_with_bf16x3_and_report_range_overflow(aliked_state, monkeypatch):
    """block1.conv2 / block2.conv1 / block2.conv2 run on two fp16 planes by default (three cross products), on three
    bf16 planes with B2S_ALIKED_CONV_NP=3: same keypoints, descriptors within fp32 rounding.  Activations beyond the
    fp16 range must not produce features: the extraction raises and names the switch."""
    from b200slam import _lib, frontend
    img = synth.frame(5, 240, 320)
    det2 = frontend.ALIKED(max_num_keypoints=512, weights=aliked_state, device="cuda:0")
    monkeypatch.setenv("B2S_ALIKED_CONV_NP", "3")
    det3 = frontend.ALIKED(max_num_keypoints=512, weights=aliked_state, device="cuda:0")
    monkeypatch.delenv("B2S_ALIKED_CONV_NP")
    f2, f3 = det2.extract_bgr(img), det3.extract_bgr(img)
    k2, k3 = f2["keypoints"][0].cpu().numpy(), f3["keypoints"][0].cpu().numpy()
    s2, s3 = set(map(tuple, np.rint(k2 * 8).astype(int).tolist())), set(map(tuple, np.rint(k3 * 8).astype(int).tolist()))
    assert len(s2 & s3) >= 0.99 * len(s3) and len(s3) > 100
    x2a, x2b = det2.debug("x2"), det3.debug("x2")
    assert rel_err(x2a, x2b) < 2e-6
    # blow the first layer up: |activation| ~ 1e7 > 65504
    big = {k: (v * 1e7 if k == "block1.conv1.weight" else v) for k, v in aliked_state.items()}
    bad = frontend.ALIKED(max_num_keypoints=512, weights=big, device="cuda:0")
    with pytest.raises(_lib.B2SError, match="B2S_ALIKED_CONV_NP=3"):
        bad.extract_bgr(img)
    with pytest.raises(_lib.B2SError, match="B2S_ALIKED_CONV_NP=3"):
        bad.extract_host(img)
    monkeypatch.setenv("B2S_ALIKED_CONV_NP", "3")
    ok = frontend.ALIKED(max_num_keypoints=512, weights=big, device="cuda:0")
    monkeypatch.delenv("B2S_ALIKED_CONV_NP")
    ok.extract_bgr(img)        # three bf16 planes keep fp32's exponent range: no error
