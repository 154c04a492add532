"""CPU: host-side logic - weight blob round trip, C-ABI surface, loud failure without a GPU,
frame generator determinism, the reference adapter running unchanged over oracle objects."""
import ctypes as C
import importlib
import os
import re
import struct
import sys
import types
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from b200slam import _lib
    hdr = open(os.path.join(ROOT, "include", "b200slam.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b2s_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in include/b200slam.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert _lib.lib.b2s_version() == 100


def test_create_fails_loudly_without_gpu(aliked_state, lightglue_state):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from b200slam import _lib, weights, frontend
    blob = weights.pack_state(aliked_state)
    cfg = _lib.AlikedCfg(); _lib.lib.b2s_aliked_default_cfg(C.byref(cfg))
    h = C.c_void_p()
    rc = _lib.lib.b2s_aliked_create(C.byref(cfg), blob, len(blob), 0, C.byref(h))
    assert rc == -4 and b"no CPU fallback" in _lib.lib.b2s_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        frontend.LightGlue(weights=lightglue_state)
    with pytest.raises(RuntimeError):
        from b200slam import features_utils as fu
        fu.init_feature_pipeline(SimpleNamespace(use_lightglue=True))


def test_weight_blob_layout_and_bad_blob(aliked_state):
    from b200slam import _lib, weights
    blob = weights.pack_state(aliked_state)
    magic, ver, n, _ = struct.unpack_from("<4sIII", blob, 0)
    assert magic == b"B2SW" and ver == 1 and n == len([k for k in aliked_state if not k.endswith("num_batches_tracked")])
    name, ndim, d0, d1, d2, d3, _, off = struct.unpack_from("<96sI4IIQ", blob, 16)
    assert name.rstrip(b"\0") == b"block1.conv1.weight" and (ndim, d0, d1, d2, d3) == (4, 16, 3, 3, 3) and off % 16 == 0
    got = np.frombuffer(blob, np.float32, 16 * 27, off)
    assert np.array_equal(got, aliked_state["block1.conv1.weight"].numpy().ravel())
    cfg = _lib.AlikedCfg(); _lib.lib.b2s_aliked_default_cfg(C.byref(cfg))
    assert (cfg.max_kp, cfg.nms_radius, cfg.resize_long) == (2048, 2, 1024) and abs(cfg.det_thresh - 0.2) < 1e-7
    lg = _lib.LgCfg(); _lib.lib.b2s_lg_default_cfg(C.byref(lg))
    assert (lg.n_layers, lg.heads, lg.dim, lg.in_dim, lg.pruning_min_kpts) == (9, 4, 256, 128, -1)


def test_state_dict_layouts_load_strict(aliked_state, lightglue_state):
    import oracle
    oracle.ALIKED().load_state_dict(aliked_state, strict=True)
    from b200slam import weights
    oracle.ALIKED(model_name="aliked-n32").load_state_dict(weights.synthetic_aliked_state("aliked-n32"), strict=True)
    missing, unexpected = oracle.LightGlue().load_state_dict(lightglue_state, strict=False)
    assert not unexpected and set(missing) <= {"confidence_thresholds"}
    legacy = {k.replace("transformers.0.self_attn", "self_attn.0"): v for k, v in lightglue_state.items()}
    assert set(weights._rename_legacy_lightglue(legacy)) == set(lightglue_state)


def test_synth_frames_are_deterministic_and_overlapping():
    from b200slam import synth
    a, b = synth.frame(0), synth.frame(1)
    assert a.shape == (376, 1241, 3) and a.dtype == np.uint8
    assert np.array_equal(a, synth.frame(0))
    assert np.array_equal(a[:, 4:], b[:, :-4]) or np.array_equal(a[1:, 4:], b[:-1, :-4]) or np.array_equal(a[:-1, 4:], b[1:, :-4])
    assert a.std() > 30


@pytest.mark.skipif(not os.path.exists("/root/reference/slam/core/features_utils.py"), reason="reference tree not present (GPU box)")
def test_reference_adapter_runs_unchanged_over_oracle_objects(aliked_state, lightglue_state):
    """Import the reference's own slam/core/features_utils.py with `lightglue` resolved to the oracle
    objects: proves the oracle honours exactly the surface the reference touches, and that
    oracle/features_utils.py restates the adapter faithfully (same outputs)."""
    import oracle
    from oracle import features_utils as ofu
    from b200slam import synth
    fake = types.ModuleType("lightglue"); fake.ALIKED, fake.LightGlue = oracle.ALIKED, oracle.LightGlue
    fu_mod = types.ModuleType("lightglue.utils"); fu_mod.rbd = oracle.rbd; fu_mod.load_image = lambda *a, **k: None
    sys.modules["lightglue"], sys.modules["lightglue.utils"] = fake, fu_mod
    try:
        spec = importlib.util.spec_from_file_location("ref_features_utils", "/root/reference/slam/core/features_utils.py")
        ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
        args = SimpleNamespace(use_lightglue=True, max_features=256, min_conf=0.5)
        det, mat = ref.init_feature_pipeline(args)
        det.load_state_dict(aliked_state, strict=True); mat.load_state_dict(lightglue_state, strict=False)
        i0, i1 = synth.frame(0, 160, 240), synth.frame(1, 160, 240)
        r0, r1 = ref.feature_extractor(args, i0, det), ref.feature_extractor(args, i1, det)
        o0, o1 = ofu.feature_extractor(args, i0, det), ofu.feature_extractor(args, i1, det)
        assert [k.pt for k in r0[0]] == [k.pt for k in o0[0]] and np.array_equal(r0[1], o0[1])
        mr = ref.feature_matcher(args, r0[0], r1[0], r0[1], r1[1], mat)
        mo = ofu.feature_matcher(args, o0[0], o1[0], o0[1], o1[1], mat)
        assert [(m.queryIdx, m.trainIdx) for m in mr] == [(m.queryIdx, m.trainIdx) for m in mo] and len(mr) > 10
    finally:
        sys.modules.pop("lightglue", None); sys.modules.pop("lightglue.utils", None)


def test_array_native_containers_quack_like_cv2_lists():
    """ndarray-backed sequences expose what the reference's consumers read (.pt / .queryIdx / .trainIdx)."""
    import cv2
    from b200slam.containers import DMatchArray, KeyPointArray
    pts = np.random.default_rng(0).random((50, 2)).astype(np.float32) * 100
    ka = KeyPointArray(pts)
    cv = ka.to_cv()
    assert len(ka) == 50 and isinstance(cv[0], cv2.KeyPoint)
    for a, b in zip(ka, cv):
        assert a.pt == b.pt and a.size == b.size and a.angle == b.angle and a.response == b.response
    assert ka[7].pt == cv[7].pt and len(ka[10:20]) == 10 and ka[10:20][0].pt == cv[10].pt
    pairs = np.stack([np.arange(20), np.arange(20)[::-1]], 1)
    ma = DMatchArray(pairs)
    mc = ma.to_cv()
    assert [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in ma] == [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in mc]
    assert np.float32([ka[m.queryIdx].pt for m in ma]).tolist() == ka.pts[ma.queryIdx].tolist()     # the reference's access pattern
    assert len(ma[np.arange(20) % 2 == 0]) == 10
