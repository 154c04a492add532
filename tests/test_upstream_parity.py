"""Auto-pin of the oracle against the REAL upstream package (cvg/LightGlue, `lightglue==0.0` of the reference's
requirements.txt:1 / setup.py:15) - the dependency `/root/reference/slam/core/features_utils.py:8-9` imports.

Neither the package nor its checkpoints exist in this offline image, so every test here SKIPS today (with the reason);
the day `import lightglue` works (and/or `~/.cache/torch/hub/checkpoints/{aliked-n16,aliked_lightglue*}.pth` appear)
they run and compare the oracle with upstream stage by stage - which would turn "parity unpinned" (oracle/__init__.py,
DESIGN.md) into a pinned oracle without touching the product.  Upstream constructors download their weights through
torch.hub; when the files are missing the download hook is replaced by the seeded synthetic state dicts, so upstream's
CODE is still exercised against the oracle's."""
import os

import numpy as np
import pytest
import torch

import oracle
from b200slam import synth, weights

lightglue = pytest.importorskip("lightglue", reason="upstream cvg/LightGlue is not installed in this image (no network): oracle "
                                                    "stays pinned to HF LightGlue / HF SuperPoint / torchvision only")


def _have(fn):
    return os.path.exists(os.path.join(weights.CKPT_DIR, fn))


@pytest.fixture()
def hub_or_synthetic(monkeypatch):
    """Real checkpoints if they are in the torch.hub cache, else feed upstream's loader the synthetic state dicts."""
    state = {"aliked": None, "lightglue": None}
    if _have("aliked-n16.pth") and (_have("aliked_lightglue_v0-1_arxiv.pth") or _have("aliked_lightglue.pth")):
        state["aliked"], _ = weights.load_aliked_state()
        state["lightglue"], _ = weights.load_lightglue_state()
        return state
    state["aliked"] = weights.synthetic_aliked_state()
    state["lightglue"] = weights.synthetic_lightglue_state()

    def fake(url, *a, **k):
        return state["aliked"] if "aliked-n" in url or "ALIKED" in url else state["lightglue"]
    monkeypatch.setattr(torch.hub, "load_state_dict_from_url", fake)
    return state


def test_preprocess_matches_upstream_image_preprocessor():
    pytest.importorskip("kornia", reason="kornia (upstream's resize backend) is not installed")
    from lightglue.utils import ImagePreprocessor
    for (h, w) in ((376, 1241), (480, 640), (1080, 1920)):
        img = oracle.bgr_to_tensor(synth.frame(1, h, w))
        up, scales = ImagePreprocessor(resize=1024, side="long", interpolation="bilinear", align_corners=None, antialias=True)(img)
        mine, sc = oracle.preprocess.resize_long_side(img, 1024)
        assert up.shape == mine.shape and torch.allclose(scales.reshape(-1), torch.as_tensor(sc, dtype=torch.float32).reshape(-1))
        assert (up - mine).abs().max() < 1e-6, (h, w)


def test_aliked_oracle_matches_upstream(hub_or_synthetic):
    pytest.importorskip("kornia", reason="kornia (upstream's resize backend) is not installed")
    up = lightglue.ALIKED(max_num_keypoints=1024).eval()
    ora = oracle.ALIKED(max_num_keypoints=1024).eval()
    ora.load_state_dict(hub_or_synthetic["aliked"], strict=True)
    img = oracle.bgr_to_tensor(synth.frame(3, 376, 1241))
    with torch.inference_mode():
        fu, fo = up.extract(img), ora.extract(img)
    ku, ko = fu["keypoints"][0].numpy(), fo["keypoints"][0].numpy()
    assert ku.shape == ko.shape
    key = lambda a: {tuple(r) for r in np.rint(a * 16).astype(np.int64).tolist()}   # noqa: E731
    assert key(ku) == key(ko), "keypoint sets differ between upstream ALIKED and the oracle"
    order = {tuple(r): i for i, r in enumerate(np.rint(ko * 16).astype(np.int64).tolist())}
    idx = np.array([order[tuple(r)] for r in np.rint(ku * 16).astype(np.int64).tolist()])
    assert np.abs(fu["descriptors"][0].numpy() - fo["descriptors"][0].numpy()[idx]).max() < 1e-5
    assert np.abs(fu["keypoint_scores"][0].numpy() - fo["keypoint_scores"][0].numpy()[idx]).max() < 1e-5


def test_lightglue_oracle_matches_upstream(hub_or_synthetic):
    from helpers import noisy_copy_pair
    up = lightglue.LightGlue(features="aliked").eval()
    ora = oracle.LightGlue(features="aliked").eval()
    ora.load_state_dict(hub_or_synthetic["lightglue"], strict=False)
    for (m, n, seed) in ((700, 512, 2), (2048, 2048, 1)):
        k0, d0, k1, d1, _ = noisy_copy_pair(m, n, seed=seed)
        data = {"image0": {"keypoints": k0[None], "descriptors": d0[None]}, "image1": {"keypoints": k1[None], "descriptors": d1[None]}}
        with torch.inference_mode():
            ru, ro = up(data), ora(data)
        assert ru["stop"] == ro["stop"]
        assert torch.equal(ru["matches"][0], ro["matches"][0]), (m, n)
        assert torch.allclose(ru["scores"][0], ro["scores"][0], rtol=1e-4, atol=1e-6)
        for k in ("matches0", "matches1", "prune0", "prune1"):
            assert torch.equal(ru[k].long(), ro[k].long()), k


def test_reference_adapter_runs_on_upstream_and_agrees_with_oracle_adapter(hub_or_synthetic):
    """The reference's own module (imported from /root/reference when present) over upstream == the oracle adapter."""
    pytest.importorskip("kornia", reason="kornia (upstream's resize backend) is not installed")
    import importlib.util
    import sys
    from types import SimpleNamespace
    path = "/root/reference/slam/core/features_utils.py"
    if not os.path.exists(path):
        pytest.skip("the reference checkout is not on this machine")
    spec = importlib.util.spec_from_file_location("ref_features_utils", path)
    ref = importlib.util.module_from_spec(spec)
    sys.modules["ref_features_utils"] = ref
    spec.loader.exec_module(ref)
    from oracle import features_utils as ofu
    args = SimpleNamespace(use_lightglue=True, detector=None, matcher=None, max_features=1024, min_conf=0.7)
    ref.torch.cuda.is_available = lambda: False            # the comparison is the CPU path
    det, mat = ref.init_feature_pipeline(args)
    odet, omat = ofu.init_feature_pipeline(args, hub_or_synthetic["aliked"], hub_or_synthetic["lightglue"])
    i0, i1 = synth.frame(10, 376, 1241), synth.frame(11, 376, 1241)
    r0, r1 = ref.feature_extractor(args, i0, det), ref.feature_extractor(args, i1, det)
    o0, o1 = ofu.feature_extractor(args, i0, odet), ofu.feature_extractor(args, i1, odet)
    mr = ref.feature_matcher(args, r0[0], r1[0], r0[1], r1[1], mat)
    mo = ofu.feature_matcher(args, o0[0], o1[0], o0[1], o1[1], omat)
    pos = lambda kps, i: tuple(np.rint(np.array(kps[i].pt) * 8).astype(int))   # noqa: E731
    assert {(pos(r0[0], m.queryIdx), pos(r1[0], m.trainIdx)) for m in mr} == {(pos(o0[0], m.queryIdx), pos(o1[0], m.trainIdx)) for m in mo}


def test_real_checkpoints_load_strict_into_the_oracle_and_the_blob():
    if not _have("aliked-n16.pth"):
        pytest.skip("no upstream checkpoints in ~/.cache/torch/hub/checkpoints")
    sa, _ = weights.load_aliked_state()
    oracle.ALIKED().load_state_dict(sa, strict=True)
    assert len(weights.pack_state(sa)) > 2_000_000
    sl, _ = weights.load_lightglue_state()
    missing, unexpected = oracle.LightGlue().load_state_dict(sl, strict=False)
    assert not [k for k in missing if "confidence_thresholds" not in k] and not unexpected
