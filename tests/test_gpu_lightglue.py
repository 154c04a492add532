"""GPU parity: LightGlue matcher (C-ABI b2s_lightglue_match[_host]) vs the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star, fp32 path): identical match index pairs,
match scores within 1e-3 relative; per-layer descriptors within 1e-4 of the oracle."""
import numpy as np
import pytest
import torch

import oracle
from b200slam import weights
from helpers import noisy_copy_pair, match_set, rel_err

pytestmark = pytest.mark.gpu

ADAPT = dict(token_bias=1.4, token_gain=6.0, match_bias=-5.0, match_gain=4.0)
ADAPT_NOSTOP = dict(token_bias=1.0, token_gain=6.0, match_bias=-5.0, match_gain=4.0)
CASES = {
    "small_ragged": (256, 200, {}, {}),
    "headline_2048": (2048, 2048, {}, {}),
    "ragged_big": (1500, 2048, {}, {}),
    "adaptive_prune_and_stop": (1024, 900, ADAPT, dict(filter_threshold=1e-6)),
    "adaptive_prune_full_depth": (512, 512, ADAPT_NOSTOP, dict(filter_threshold=1e-6)),
    "adaptivity_off": (512, 512, {}, dict(depth_confidence=-1, width_confidence=-1)),
    "cuda_flash_pruning_threshold": (2048, 1800, ADAPT_NOSTOP, dict(filter_threshold=1e-6, pruning_threshold=1536)),
}


def _models(wkw, mkw, n_layers=9):
    from b200slam import frontend
    sd = weights.synthetic_lightglue_state(seed=0, n_layers=n_layers, **wkw)
    ora = oracle.LightGlue(n_layers=n_layers, **{k: v for k, v in mkw.items() if k not in ("precision", "max_kp")}).eval()
    ora.load_state_dict(sd, strict=False)
    return ora, frontend.LightGlue(weights=sd, **mkw)


def _oracle_run(ora, k0, d0, k1, d1, size=None):
    i0 = {"keypoints": k0[None], "descriptors": d0[None]}
    i1 = {"keypoints": k1[None], "descriptors": d1[None]}
    if size is not None:
        i0["image_size"] = torch.tensor([size], dtype=torch.float32); i1["image_size"] = i0["image_size"]
    return ora({"image0": i0, "image1": i1})


@pytest.mark.parametrize("name", list(CASES))
def test_match_parity_through_c_abi(name):
    m, n, wkw, mkw = CASES[name]
    ora, mat = _models(wkw, mkw)
    k0, d0, k1, d1, _ = noisy_copy_pair(m, n, seed=1)
    ro = _oracle_run(ora, k0, d0, k1, d1)
    rg = mat.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy(), full=True)
    mo = ro["matches"][0].numpy()
    assert len(mo) > 20, "vacuous case"
    assert rg["stop"] == ro["stop"]
    assert np.array_equal(rg["matches"], mo), f"match pairs differ: {len(match_set(mo) ^ match_set(rg['matches']))} of {len(mo)}"
    assert rel_err(rg["scores"], ro["scores"][0].numpy()) < 1e-3          # tolerance: 1e-3 relative
    for k in ("matches0", "matches1", "prune0", "prune1"):
        assert np.array_equal(rg[k], ro[k][0].long().numpy()), k
    for k in ("matching_scores0", "matching_scores1"):
        assert rel_err(rg[k], ro[k][0].numpy()) < 1e-3, k
    if name.startswith("adaptive"):
        assert ro["prune0"][0].min() < ro["stop"], "pruning must actually happen in this case"


def test_config4_matcher_4096_full_depth_vs_oracle():
    """BASELINE config 4's matcher: 4096 x 4096 keypoints, depth_confidence = width_confidence = -1 (all 9 layers, no
    pruning), fp32 path against the CPU oracle: identical match index pairs, scores within 1e-3 relative."""
    mkw = dict(depth_confidence=-1, width_confidence=-1)
    ora, mat = _models({}, dict(mkw, max_kp=4096))
    k0, d0, k1, d1, _ = noisy_copy_pair(4096, 4096, seed=7)
    k0 = k0 * torch.tensor([1920.0 / 1241.0, 1080.0 / 376.0]); k1 = k1 * torch.tensor([1920.0 / 1241.0, 1080.0 / 376.0])
    ro = _oracle_run(ora, k0, d0, k1, d1)
    rg = mat.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy(), full=True)
    mo = ro["matches"][0].numpy()
    assert len(mo) > 500 and rg["stop"] == ro["stop"] == 9
    assert np.array_equal(rg["matches"], mo), f"match pairs differ: {len(match_set(mo) ^ match_set(rg['matches']))} of {len(mo)}"
    assert rel_err(rg["scores"], ro["scores"][0].numpy()) < 1e-3          # tolerance: 1e-3 relative
    assert np.array_equal(rg["matches0"], ro["matches0"][0].numpy()) and np.array_equal(rg["matches1"], ro["matches1"][0].numpy())
    assert (rg["prune0"] == 9).all() and (rg["prune1"] == 9).all()


def _window_feats(n_kf, lo, hi, seed=100):
    g = torch.Generator().manual_seed(seed)
    base = noisy_copy_pair(hi, hi, seed=seed)
    feats = []
    for f in range(n_kf):
        n = int(torch.randint(lo, hi + 1, (1,), generator=g))
        sel = torch.randperm(hi, generator=g)[:n]
        d = torch.nn.functional.normalize(base[1][sel] + 0.03 * torch.randn(n, 128, generator=g), dim=1)
        k = base[0][sel] + 2.0 * f + torch.randn(n, 2, generator=g)
        feats.append((k.contiguous(), d.contiguous()))
    return feats


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batched_window_equals_single_pair_calls(precision):
    """BASELINE config 3 shape: 16 keyframes, all 120 pairs in batched launch sequences (the pair is a grid dimension of
    every kernel) == 120 single-pair calls, bit for bit (matches, scores, counts, executed layers), ragged sizes, with
    weights that make pairs prune and exit at different layers."""
    from b200slam import frontend, sharding
    sd = weights.synthetic_lightglue_state(seed=0, **ADAPT)
    mat = frontend.LightGlue(weights=sd, precision=precision, filter_threshold=1e-6, max_kp=640)
    feats = [(k.cuda(), d.cuda()) for k, d in _window_feats(16, 380, 640)]
    pairs = sharding.window_pairs(16)
    single = {}
    for (i, j) in pairs:
        r = mat.match_device(feats[i][0], feats[i][1], feats[j][0], feats[j][1], full=False)
        torch.cuda.synchronize()
        n = int(r["n"])
        single[(i, j)] = (r["matches"][:n].cpu().numpy(), r["scores"][:n].cpu().numpy(), int(r["stop"]))
    stops = {v[2] for v in single.values()}
    assert len(stops) > 1, f"the pairs must not all stop at the same layer ({stops})"
    for max_batch in (0, 5):
        out = mat.match_batch_device([f[0] for f in feats], [f[1] for f in feats], pairs, max_batch=max_batch)
        torch.cuda.synchronize()
        nb, mb, sb, stb = out["n"].cpu().numpy(), out["matches"].cpu().numpy(), out["scores"].cpu().numpy(), out["stop"].cpu().numpy()
        for p, key in enumerate(pairs):
            ms, ss, st = single[key]
            assert nb[p] == len(ms) and stb[p] == st, (key, nb[p], len(ms), stb[p], st)
            assert np.array_equal(mb[p, :nb[p]], ms) and np.array_equal(sb[p, :nb[p]], ss), key
    assert sum(len(v[0]) for v in single.values()) > 120 * 10, "vacuous"


def test_batch_vs_oracle_headline_size():
    """A batch of 3 pairs at 2048 keypoints against the CPU oracle pair by pair: identical match index pairs."""
    ora, mat = _models({}, dict(max_kp=2048))
    feats = _window_feats(3, 1800, 2048, seed=5)
    pairs = [(0, 1), (0, 2), (1, 2)]
    out = mat.match_batch_device([f[0].cuda() for f in feats], [f[1].cuda() for f in feats], pairs)
    torch.cuda.synchronize()
    for p, (i, j) in enumerate(pairs):
        ro = _oracle_run(ora, feats[i][0], feats[i][1], feats[j][0], feats[j][1])
        n = int(out["n"][p])
        assert len(ro["matches"][0]) > 50
        assert np.array_equal(out["matches"][p, :n].cpu().numpy(), ro["matches"][0].numpy()), (i, j)
        assert rel_err(out["scores"][p, :n].cpu().numpy(), ro["scores"][0].numpy()) < 1e-3
        assert int(out["stop"][p]) == ro["stop"]


def test_batch_edge_cases():
    """Empty frames inside a batch, a single pair, more pairs than one launch sequence holds."""
    ora, mat = _models({}, {})
    feats = _window_feats(4, 60, 130, seed=9)
    feats.append((torch.zeros(0, 2), torch.zeros(0, 128)))
    kl, dl = [f[0].cuda() for f in feats], [f[1].cuda() for f in feats]
    pairs = [(0, 4), (4, 1), (0, 1), (2, 3)] + [(a, b) for a in range(4) for b in range(4) if a != b] * 2
    assert len(pairs) > mat.max_batch
    out = mat.match_batch_device(kl, dl, pairs)
    torch.cuda.synchronize()
    n = out["n"].cpu().numpy()
    assert n[0] == 0 and n[1] == 0
    for p, (i, j) in enumerate(pairs):
        if 4 in (i, j):
            continue
        ro = _oracle_run(ora, feats[i][0], feats[i][1], feats[j][0], feats[j][1])
        assert np.array_equal(out["matches"][p, :n[p]].cpu().numpy(), ro["matches"][0].numpy()), (p, i, j)
    assert mat.workspace_bytes(2048, 4) > mat.workspace_bytes(2048, 1) > 50e6


def test_layer_taps_and_similarity():
    ora, mat = _models({}, {})
    ora.record_taps = True
    mat.set_debug(True)
    k0, d0, k1, d1, _ = noisy_copy_pair(700, 512, seed=2)
    _oracle_run(ora, k0, d0, k1, d1)
    mat.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy())
    for i, (a, b) in enumerate(ora.taps["layers"]):
        assert rel_err(mat.debug(f"layer{i}_0").reshape(-1, 256), a[0].numpy()) < 1e-4, f"layer {i} image0"
        assert rel_err(mat.debug(f"layer{i}_1").reshape(-1, 256), b[0].numpy()) < 1e-4, f"layer {i} image1"
    assert rel_err(mat.debug("sim").reshape(700, 512), ora.taps["sim"][0].numpy()) < 1e-4


def test_image_size_path_and_torch_call_contract():
    """The upstream __call__ contract used at features_utils.py:157-161 / :237 (dict of tensors)."""
    ora, mat = _models({}, {})
    k0, d0, k1, d1, _ = noisy_copy_pair(300, 300, seed=3)
    ro = _oracle_run(ora, k0, d0, k1, d1, size=(1241.0, 376.0))
    sz = torch.tensor([[1241.0, 376.0]])
    rg = mat({"image0": {"keypoints": k0[None].cuda(), "descriptors": d0[None].cuda(), "image_size": sz},
              "image1": {"keypoints": k1[None].cuda(), "descriptors": d1[None].cuda(), "image_size": sz}})
    assert torch.equal(rg["matches"][0].cpu(), ro["matches"][0])
    assert rg["matches"][0].dtype == torch.int64 and rg["matches0"].shape == (1, 300)
    assert rg["stop"] == ro["stop"] and torch.equal(rg["prune0"].cpu(), ro["prune0"])
    assert next(mat.parameters()).device.type == "cuda"          # features_utils.py:131
    assert mat.eval() is mat and mat.to("cuda") is mat


@pytest.mark.parametrize("m,n", [(0, 10), (10, 0), (1, 1), (5, 3), (33, 65)])
def test_edge_sizes(m, n):
    ora, mat = _models({}, {})
    k0, d0, k1, d1, _ = noisy_copy_pair(max(m, 1), max(n, 1), seed=4)
    k0, d0, k1, d1 = k0[:m], d0[:m], k1[:n], d1[:n]
    rg = mat.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy(), full=True)
    if m == 0 or n == 0:
        assert len(rg["matches"]) == 0 and (rg["matches0"] == -1).all() and (rg["matches1"] == -1).all()
        return
    ro = _oracle_run(ora, k0, d0, k1, d1)
    assert np.array_equal(rg["matches"], ro["matches"][0].numpy())
    assert np.array_equal(rg["matches0"], ro["matches0"][0].numpy())


def test_device_api_equals_host_api_and_batch():
    from b200slam import _lib
    import ctypes as C
    ora, mat = _models({}, {})
    feats = [noisy_copy_pair(400 + 50 * i, 400 + 50 * i, seed=10 + i)[:2] for i in range(3)]
    ref = {}
    for (a, b) in [(0, 1), (0, 2), (1, 2)]:
        ref[(a, b)] = mat.match_host(feats[a][0].numpy(), feats[a][1].numpy(), feats[b][0].numpy(), feats[b][1].numpy())["matches"]
        dv = mat.match_device(feats[a][0].cuda(), feats[a][1].cuda(), feats[b][0].cuda(), feats[b][1].cuda(), full=False)
        torch.cuda.synchronize()
        assert np.array_equal(dv["matches"][: int(dv["n"])].cpu().numpy(), ref[(a, b)])
    # b2s_lightglue_match_batch over the packed window (BASELINE config 3 shape, small)
    kp = torch.cat([f[0] for f in feats]).cuda(); de = torch.cat([f[1] for f in feats]).cuda()
    cu = np.cumsum([0] + [len(f[0]) for f in feats]).astype(np.int32)
    pi, pj = np.array([0, 0, 1], np.int32), np.array([1, 2, 2], np.int32)
    stride = 512
    mt = torch.zeros((3, stride, 2), dtype=torch.int32, device="cuda"); sc = torch.zeros((3, stride), device="cuda")
    nm = torch.zeros(3, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib.b2s_lightglue_match_batch(mat._handle, kp.data_ptr(), de.data_ptr(), cu.ctypes.data, 3, pi.ctypes.data,
                                                  pj.ctypes.data, 3, torch.cuda.current_stream().cuda_stream, stride,
                                                  mt.data_ptr(), sc.data_ptr(), nm.data_ptr()), "batch")
    torch.cuda.synchronize()
    for p, key in enumerate([(0, 1), (0, 2), (1, 2)]):
        assert np.array_equal(mt[p, : int(nm[p])].cpu().numpy(), ref[key])


def test_golden_fixture_matches():
    """Committed oracle fixture (tests/golden): GPU matcher on the stored features reproduces the stored matches."""
    import json, os
    g = os.path.join(os.path.dirname(__file__), "golden")
    meta = json.load(open(os.path.join(g, "meta.json"))); z = np.load(os.path.join(g, "aliked_lg_small.npz"))
    from b200slam import frontend
    mat = frontend.LightGlue(weights=weights.synthetic_lightglue_state(meta["seed"]))
    size = (float(meta["W"]), float(meta["H"]))
    rg = mat.match_host(z["kp0"], z["desc0"], z["kp1"], z["desc1"], size0=size, size1=size)
    assert np.array_equal(rg["matches"], z["matches"]) and rg["stop"] == int(z["stop"])
    assert rel_err(rg["scores"], z["scores"]) < 1e-3
