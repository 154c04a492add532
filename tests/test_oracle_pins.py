"""CPU: pin the oracle against independent implementations available in the image.
The reference itself holds no golden vectors for this path (parity unpinned);
these cross-checks are the strongest anchors available offline:
  * HuggingFace `transformers` LightGlue port  <-> oracle.lightglue (same weights, same inputs)
  * HF SuperPoint simple_nms                   <-> oracle.aliked.simple_nms
  * torchvision.ops.deform_conv2d probes       <-> the offset-channel convention we restate
  * committed golden fixtures (tests/golden)   <-> oracle outputs today
"""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import lightglue as olg
from b200slam import weights, synth
from helpers import noisy_copy_pair, match_set

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hf_layers(sd, n_layers):
    from transformers.models.lightglue.configuration_lightglue import LightGlueConfig
    from transformers.models.lightglue import modeling_lightglue as hf
    cfg = LightGlueConfig(descriptor_dim=256, num_hidden_layers=n_layers, num_attention_heads=4)
    cfg._attn_implementation = "eager"
    layers = [hf.LightGlueTransformerLayer(cfg, i).eval() for i in range(n_layers)]
    assign = [hf.LightGlueMatchAssignmentLayer(cfg).eval() for _ in range(n_layers)]
    posenc = hf.LightGluePositionalEncoder(cfg).eval()
    posenc.projector.weight.data.copy_(sd["posenc.Wr.weight"])
    for i, (ly, asg) in enumerate(zip(layers, assign)):
        p = f"transformers.{i}"
        wqkv, bqkv = sd[f"{p}.self_attn.Wqkv.weight"], sd[f"{p}.self_attn.Wqkv.bias"]
        # upstream Wqkv rows are (head, dim, {q,k,v}) interleaved
        w = wqkv.view(4, 64, 3, 256); b = bqkv.view(4, 64, 3)
        for j, proj in enumerate((ly.self_attention.q_proj, ly.self_attention.k_proj, ly.self_attention.v_proj)):
            proj.weight.data.copy_(w[:, :, j].reshape(256, 256)); proj.bias.data.copy_(b[:, :, j].reshape(256))
        ly.self_attention.o_proj.weight.data.copy_(sd[f"{p}.self_attn.out_proj.weight"])
        ly.self_attention.o_proj.bias.data.copy_(sd[f"{p}.self_attn.out_proj.bias"])
        for proj, nm in ((ly.cross_attention.q_proj, "to_qk"), (ly.cross_attention.k_proj, "to_qk"),
                         (ly.cross_attention.v_proj, "to_v"), (ly.cross_attention.o_proj, "to_out")):
            proj.weight.data.copy_(sd[f"{p}.cross_attn.{nm}.weight"]); proj.bias.data.copy_(sd[f"{p}.cross_attn.{nm}.bias"])
        for mlp, blk in ((ly.self_mlp, "self_attn"), (ly.cross_mlp, "cross_attn")):
            mlp.fc1.weight.data.copy_(sd[f"{p}.{blk}.ffn.0.weight"]); mlp.fc1.bias.data.copy_(sd[f"{p}.{blk}.ffn.0.bias"])
            mlp.layer_norm.weight.data.copy_(sd[f"{p}.{blk}.ffn.1.weight"]); mlp.layer_norm.bias.data.copy_(sd[f"{p}.{blk}.ffn.1.bias"])
            mlp.fc2.weight.data.copy_(sd[f"{p}.{blk}.ffn.3.weight"]); mlp.fc2.bias.data.copy_(sd[f"{p}.{blk}.ffn.3.bias"])
        asg.final_projection.weight.data.copy_(sd[f"log_assignment.{i}.final_proj.weight"])
        asg.final_projection.bias.data.copy_(sd[f"log_assignment.{i}.final_proj.bias"])
        asg.matchability.weight.data.copy_(sd[f"log_assignment.{i}.matchability.weight"])
        asg.matchability.bias.data.copy_(sd[f"log_assignment.{i}.matchability.bias"])
    return hf, layers, assign, posenc


@torch.no_grad()
def test_lightglue_oracle_matches_hf_port():
    n_layers, n = 3, 192
    sd = weights.synthetic_lightglue_state(seed=3, n_layers=n_layers)
    k0, d0, k1, d1, _ = noisy_copy_pair(n, n, seed=1)
    ora = oracle.LightGlue(n_layers=n_layers, depth_confidence=-1, width_confidence=-1).eval()
    ora.load_state_dict(sd, strict=False)
    ora.record_taps = True
    size = torch.tensor([[1241.0, 376.0]])
    ro = ora({"image0": {"keypoints": k0[None], "descriptors": d0[None], "image_size": size},
              "image1": {"keypoints": k1[None], "descriptors": d1[None], "image_size": size}})
    hf, layers, assign, posenc = _hf_layers(sd, n_layers)
    kp = hf.normalize_keypoints(torch.stack([k0, k1]), 376, 1241)
    desc = torch.nn.functional.linear(torch.stack([d0, d1]), sd["input_proj.weight"], sd["input_proj.bias"])
    enc = posenc(kp)[0]
    for i, ly in enumerate(layers):
        desc = ly(desc, enc, attention_mask=None)[0]
        o0, o1 = ora.taps["layers"][i]
        assert torch.allclose(desc[0], o0[0], atol=2e-4, rtol=1e-4), f"layer {i} image0"
        assert torch.allclose(desc[1], o1[0], atol=2e-4, rtol=1e-4), f"layer {i} image1"
    scores = assign[-1](desc, None)
    assert torch.allclose(scores, ora.taps["log_assignment"], atol=2e-3, rtol=1e-4)
    m, ms = hf.get_matches_from_scores(scores, 0.1)
    assert torch.equal(m[0], ro["matches0"][0]) and torch.equal(m[1], ro["matches1"][0])
    assert torch.allclose(ms[0], ro["matching_scores0"][0], atol=1e-4)
    assert len(ro["matches"][0]) > n // 2, "synthetic weights must give non-vacuous matches"


def test_confidence_thresholds_match_hf_formula():
    ora = oracle.LightGlue()
    for i in range(9):
        assert float(ora.confidence_thresholds[i]) == pytest.approx(np.clip(0.8 + 0.1 * np.exp(-4.0 * i / 9), 0, 1), rel=1e-6)


def test_simple_nms_matches_hf_superpoint():
    from transformers.models.superpoint.modeling_superpoint import simple_nms as hf_nms
    g = torch.Generator().manual_seed(0)
    s = torch.rand(1, 57, 83, generator=g)
    assert torch.equal(hf_nms(s, 2), oracle.simple_nms(s[None], 2)[0])


def test_deform_conv_offset_convention():
    """torchvision probes behind SURVEY A.2: channel 2k = dy, 2k+1 = dx of tap k = ky*3+kx; zero outside."""
    from torchvision.ops import deform_conv2d
    x = torch.arange(25, dtype=torch.float32).view(1, 1, 5, 5)
    w = torch.zeros(1, 1, 3, 3); w[0, 0, 1, 1] = 1.0   # centre tap only (k = 4)
    off = torch.zeros(1, 18, 5, 5)
    off[0, 8] = 1.0    # dy of tap 4
    y = deform_conv2d(x, off, w, padding=1)
    assert torch.equal(y[0, 0, :4], x[0, 0, 1:]) and torch.all(y[0, 0, 4] == 0)
    off.zero_(); off[0, 9] = 0.5   # dx of tap 4
    y = deform_conv2d(x, off, w, padding=1)
    assert torch.allclose(y[0, 0, :, :4], x[0, 0, :, :4] + 0.5)
    assert torch.allclose(y[0, 0, :, 4], x[0, 0, :, 4] * 0.5)


def test_resized_shapes_of_baseline_configs():
    from oracle.preprocess import resized_shape, blur_params
    assert resized_shape(376, 1241) == (310, 1024)
    assert resized_shape(480, 640) == (768, 1024)
    assert resized_shape(1080, 1920) == (576, 1024)
    assert blur_params(480, 640, 768, 1024) is None
    ky, kx, sy, sx = blur_params(1080, 1920, 576, 1024)
    assert (ky, kx) == (3, 3) and sy == pytest.approx(0.4375)


@torch.no_grad()
def test_oracle_against_committed_golden():
    """Golden fixtures were produced by tests/golden/make_golden.py from this oracle; they freeze
    its behaviour so that later edits cannot drift silently (and travel to the GPU box)."""
    meta = json.load(open(os.path.join(GOLD, "meta.json")))
    z = np.load(os.path.join(GOLD, "aliked_lg_small.npz"))
    sd_a = weights.synthetic_aliked_state(meta["model"], meta["seed"])
    det = oracle.ALIKED(model_name=meta["model"], max_num_keypoints=meta["max_kp"]).eval()
    det.load_state_dict(sd_a, strict=True)
    H, W = meta["H"], meta["W"]
    f0 = det.extract(oracle.bgr_to_tensor(synth.frame(0, H, W)))
    f1 = det.extract(oracle.bgr_to_tensor(synth.frame(1, H, W)))
    np.testing.assert_allclose(f0["keypoints"][0].numpy(), z["kp0"], atol=1e-3)
    np.testing.assert_allclose(f0["descriptors"][0].numpy(), z["desc0"], atol=1e-4)
    mat = oracle.LightGlue().eval()
    mat.load_state_dict(weights.synthetic_lightglue_state(meta["seed"]), strict=False)
    r = mat({"image0": f0, "image1": f1})
    assert match_set(r["matches"][0]) == match_set(z["matches"])
    np.testing.assert_allclose(r["scores"][0].numpy(), z["scores"], atol=1e-3)
    assert len(z["matches"]) >= meta["min_matches"]


def test_preprocess_blur_and_resize_against_opencv():
    """Independent pin of the kornia-style anti-aliased resize restated in oracle/preprocess.py (upstream
    ImagePreprocessor: gaussian_blur2d with a reflect border, then bilinear interpolation with half-pixel centres):
    OpenCV implements the same two operations (cv2.GaussianBlur with BORDER_REFLECT_101 and an explicit sigma;
    cv2.resize INTER_LINEAR) - same kernel formula, same border rule, same sampling grid."""
    import cv2
    import torch.nn.functional as F
    from oracle import preprocess as P
    rng = np.random.default_rng(0)
    for (h, w) in [(376, 1241), (480, 640), (1080, 1920)]:
        img = rng.random((h, w, 3), dtype=np.float32)
        t = torch.from_numpy(img).permute(2, 0, 1)[None]
        hn, wn = P.resized_shape(h, w, 1024)
        bp = P.blur_params(h, w, hn, wn)
        if bp is not None:
            ky, kx, sy, sx = bp
            mine = P.gaussian_blur2d(t, ky, kx, sy, sx)[0].permute(1, 2, 0).numpy()
            theirs = cv2.GaussianBlur(img, (kx, ky), sigmaX=sx, sigmaY=sy, borderType=cv2.BORDER_REFLECT_101)
            assert np.abs(mine - theirs).max() < 2e-6, (h, w)
            # the 1-D kernel itself
            assert np.allclose(P.gaussian_kernel1d(kx, sx).numpy(), cv2.getGaussianKernel(kx, sx, cv2.CV_32F).ravel(), atol=1e-7)
            src = theirs
        else:
            src = img
        # bilinear, half-pixel centres (F.interpolate align_corners=False == cv2.INTER_LINEAR) on the SAME blurred input
        mine = F.interpolate(torch.from_numpy(src).permute(2, 0, 1)[None], size=(hn, wn), mode="bilinear", align_corners=None)[0].permute(1, 2, 0).numpy()
        theirs = cv2.resize(src, (wn, hn), interpolation=cv2.INTER_LINEAR)
        assert np.abs(mine - theirs).max() < 1e-4, (h, w)      # float32 source coordinates differ by ~1e-5 px between the two; white noise turns that into ~2e-5.  A half-pixel or align_corners mix-up would be ~1e-1
        # and the composed function the oracle uses
        full, scales = P.resize_long_side(t, 1024)
        assert np.abs(full[0].permute(1, 2, 0).numpy() - theirs).max() < 1e-4
        assert np.allclose(scales.numpy(), [wn / w, hn / h])
