"""Weights for the ALIKED + LightGlue frontend: upstream state-dict layouts (SURVEY.md
Appendix B), a seeded synthetic factory for offline runs (no network => no checkpoints),
a loader for real upstream checkpoints when they exist, and the flat blob format the
C-ABI consumes (`b2s_aliked_create` / `b2s_lightglue_create`, include/b200slam.h).

Reference call sites this replaces: `lightglue.ALIKED(...)` / `lightglue.LightGlue(...)`
constructors at `/root/reference/slam/core/features_utils.py:25-26`, which download
`aliked-n16.pth` / `aliked_lightglue.pth` through torch.hub.
"""
from __future__ import annotations

import math
import os
import warnings
import struct
from collections import OrderedDict

import numpy as np
import torch

ALIKED_CFGS = {  # c1, c2, c3, c4, dim, K, M
    "aliked-n16": (16, 32, 64, 128, 128, 3, 16),
    "aliked-n32": (16, 32, 64, 128, 128, 3, 32),
}
CKPT_DIR = os.path.expanduser("~/.cache/torch/hub/checkpoints")


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _conv(sd, g, name, cout, cin, k, bias):
    fan_in = cin * k * k
    b = 1.0 / math.sqrt(fan_in)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    sd[name + ".weight"] = _uniform(g, (cout, cin, k, k), b)
    if bias:
        sd[name + ".bias"] = _uniform(g, (cout,), b)


def _linear(sd, g, name, cout, cin, bias=True):
    b = 1.0 / math.sqrt(cin)
    sd[name + ".weight"] = _uniform(g, (cout, cin), b)
    if bias:
        sd[name + ".bias"] = _uniform(g, (cout,), b)


def _bn(sd, g, name, c):
    # mildly randomised running stats so that BN folding is actually exercised
    sd[name + ".weight"] = 1.0 + _uniform(g, (c,), 0.2)
    sd[name + ".bias"] = _uniform(g, (c,), 0.1)
    sd[name + ".running_mean"] = _uniform(g, (c,), 0.1)
    sd[name + ".running_var"] = 1.0 + _uniform(g, (c,), 0.2)
    sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def synthetic_aliked_state(model_name="aliked-n16", seed=0) -> "OrderedDict[str, torch.Tensor]":
    c1, c2, c3, c4, dim, K, M = ALIKED_CFGS[model_name]
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    _conv(sd, g, "block1.conv1", c1, 3, 3, False); _bn(sd, g, "block1.bn1", c1)
    _conv(sd, g, "block1.conv2", c1, c1, 3, False); _bn(sd, g, "block1.bn2", c1)
    _conv(sd, g, "block2.conv1", c2, c1, 3, False); _bn(sd, g, "block2.bn1", c2)
    _conv(sd, g, "block2.conv2", c2, c2, 3, False); _bn(sd, g, "block2.bn2", c2)
    _conv(sd, g, "block2.downsample", c2, c1, 1, True)
    for blk, ci, co in (("block3", c2, c3), ("block4", c3, c4)):
        _conv(sd, g, f"{blk}.conv1.offset_conv", 18, ci, 3, True)
        _conv(sd, g, f"{blk}.conv1.regular_conv", co, ci, 3, False); _bn(sd, g, f"{blk}.bn1", co)
        _conv(sd, g, f"{blk}.conv2.offset_conv", 18, co, 3, True)
        _conv(sd, g, f"{blk}.conv2.regular_conv", co, co, 3, False); _bn(sd, g, f"{blk}.bn2", co)
        _conv(sd, g, f"{blk}.downsample", co, ci, 1, True)
    for i, ci in enumerate((c1, c2, c3, c4), start=1):
        _conv(sd, g, f"conv{i}", dim // 4, ci, 1, False)
    _conv(sd, g, "score_head.0", 8, dim, 1, False)
    _conv(sd, g, "score_head.2", 4, 8, 3, False)
    _conv(sd, g, "score_head.4", 4, 4, 3, False)
    _conv(sd, g, "score_head.6", 1, 4, 3, False)
    _conv(sd, g, "desc_head.offset_conv.0", 2 * M, dim, K, True)
    _conv(sd, g, "desc_head.offset_conv.2", 2 * M, 2 * M, 1, True)
    _conv(sd, g, "desc_head.sf_conv", dim, dim, 1, False)
    sd["desc_head.agg_weights"] = _uniform(g, (M, dim, dim), 0.02)
    # Default-initialised weights are degenerate offline (scores all ~0.5, descriptors all
    # parallel, sub-0.1 px deformable offsets), which would make parity vacuous.  Rescale so
    # that: scores span ~0.1..0.8 (threshold 0.2 and the top-k cut both bite), DCN / SDDH
    # offsets are O(1) px (bilinear gathers exercised), and the aggregation cancels the
    # component shared by all sample positions (descriptors become discriminative).
    for k in ("score_head.0", "score_head.2", "score_head.4", "score_head.6"):
        sd[k + ".weight"] *= 3.0
    sd["desc_head.offset_conv.2.bias"] = _uniform(g, (2 * M,), 4.0)
    sd["desc_head.offset_conv.0.weight"] *= 8.0
    sd["desc_head.offset_conv.2.weight"] *= 2.0
    a = sd["desc_head.agg_weights"]
    sd["desc_head.agg_weights"] = a - a.mean(0, keepdim=True)
    for blk in ("block3", "block4"):
        for c in ("conv1", "conv2"):
            sd[f"{blk}.{c}.offset_conv.weight"] *= 8.0
            sd[f"{blk}.{c}.offset_conv.bias"] *= 8.0
    return sd


def synthetic_lightglue_state(seed=0, final_scale=16.0, n_layers=9, dim=256, in_dim=128,
                              token_bias=-2.0, token_gain=0.1, match_bias=4.0, match_gain=0.0):
    """Seeded weights giving non-degenerate matcher behaviour: default
    nn.Linear init, ffn[3] x0.1, final_proj = S*I, matchability/token heads biased so that
    early exit / pruning do not fire (defaults) or do (override token_*/match_*)."""
    g = torch.Generator().manual_seed(seed + 1000)
    sd = OrderedDict()
    _linear(sd, g, "input_proj", dim, in_dim)
    sd["posenc.Wr.weight"] = torch.randn((dim // 4 // 2, 2), generator=g, dtype=torch.float32)
    for i in range(n_layers):
        for blk, lins in (("self_attn", (("Wqkv", 3 * dim, dim), ("out_proj", dim, dim))),
                          ("cross_attn", (("to_qk", dim, dim), ("to_v", dim, dim), ("to_out", dim, dim)))):
            p = f"transformers.{i}.{blk}"
            for nm, co, ci in lins:
                _linear(sd, g, f"{p}.{nm}", co, ci)
            _linear(sd, g, f"{p}.ffn.0", 2 * dim, 2 * dim)
            sd[f"{p}.ffn.1.weight"] = 1.0 + _uniform(g, (2 * dim,), 0.1)
            sd[f"{p}.ffn.1.bias"] = _uniform(g, (2 * dim,), 0.05)
            _linear(sd, g, f"{p}.ffn.3", dim, 2 * dim)
            sd[f"{p}.ffn.3.weight"] *= 0.1
            sd[f"{p}.ffn.3.bias"].zero_()
        _linear(sd, g, f"log_assignment.{i}.matchability", 1, dim)
        sd[f"log_assignment.{i}.matchability.weight"] *= match_gain
        sd[f"log_assignment.{i}.matchability.bias"].fill_(match_bias)
        sd[f"log_assignment.{i}.final_proj.weight"] = final_scale * torch.eye(dim)
        sd[f"log_assignment.{i}.final_proj.bias"] = torch.zeros(dim)
        if i < n_layers - 1:
            _linear(sd, g, f"token_confidence.{i}.token.0", 1, dim)
            sd[f"token_confidence.{i}.token.0.weight"] *= token_gain
            sd[f"token_confidence.{i}.token.0.bias"].fill_(token_bias)
    return sd


def _rename_legacy_lightglue(sd):
    """upstream renames `self_attn.{i}.*`/`cross_attn.{i}.*` -> `transformers.{i}.*` on load."""
    out = OrderedDict()
    for k, v in sd.items():
        for blk in ("self_attn", "cross_attn"):
            if k.startswith(blk + "."):
                i, rest = k[len(blk) + 1:].split(".", 1)
                k = f"transformers.{i}.{blk}.{rest}"
                break
        out[k] = v
    return out


class MissingCheckpointError(FileNotFoundError):
    pass


def _synthetic_allowed(allow_synthetic) -> bool:
    """Explicit opt-in only: the argument wins, else B2S_SYNTHETIC_WEIGHTS=1 (benchmarks / tests on boxes without the
    checkpoints).  The reference always runs real weights (torch.hub download), so silently matching with random
    weights would quietly degrade a SLAM run instead of failing."""
    if allow_synthetic is not None:
        return bool(allow_synthetic)
    return os.environ.get("B2S_SYNTHETIC_WEIGHTS", "0") == "1"


def _missing(what, tried):
    return MissingCheckpointError(
        f"b200slam: no {what} checkpoint found (tried {', '.join(tried)}). Put the upstream file there, pass its path "
        "(args.aliked_weights / args.lightglue_weights or weights=...), or opt in to seeded synthetic weights with "
        "args.synthetic_weights=True / B2S_SYNTHETIC_WEIGHTS=1 (benchmarks and tests only: matches are meaningless).")


def load_aliked_state(model_name="aliked-n16", seed=0, path=None, allow_synthetic=None):
    """(state_dict, source): `path` if given, else the reference's torch.hub cache; synthetic only on opt-in."""
    tried = [path] if path else [os.path.join(CKPT_DIR, f"{model_name}.pth")]
    for p in tried:
        if os.path.exists(p):
            return torch.load(p, map_location="cpu"), p
    if path or not _synthetic_allowed(allow_synthetic):
        raise _missing(f"ALIKED ({model_name})", tried)
    warnings.warn(f"b200slam: ALIKED runs on SYNTHETIC weights (seed={seed}); keypoints are not meaningful", stacklevel=2)
    return synthetic_aliked_state(model_name, seed), f"synthetic(seed={seed})"


def load_lightglue_state(seed=0, path=None, allow_synthetic=None, **kw):
    tried = [path] if path else [os.path.join(CKPT_DIR, fn) for fn in ("aliked_lightglue_v0-1_arxiv.pth", "aliked_lightglue.pth")]
    for p in tried:
        if os.path.exists(p):
            return _rename_legacy_lightglue(torch.load(p, map_location="cpu")), p
    if path or not _synthetic_allowed(allow_synthetic):
        raise _missing("LightGlue (aliked)", tried)
    warnings.warn(f"b200slam: LightGlue runs on SYNTHETIC weights (seed={seed}); matches are not meaningful", stacklevel=2)
    return synthetic_lightglue_state(seed, **kw), f"synthetic(seed={seed})"


# ----------------------------------------------------------------------------------------
# Flat blob consumed by the C-ABI:  header | table | fp32 payload (16-byte aligned tensors)
#   header : char magic[4]="B2SW"; u32 version=1; u32 n_tensors; u32 reserved
#   entry  : char name[96]; u32 ndim; u32 dims[4]; u32 reserved; u64 byte_offset   (128 B)
# ----------------------------------------------------------------------------------------
_ENTRY = struct.Struct("<96sI4IIQ")
assert _ENTRY.size == 128


def pack_state(sd) -> bytes:
    items = [(k, v.detach().to(torch.float32).contiguous().cpu().numpy())
             for k, v in sd.items() if not k.endswith("num_batches_tracked")]
    head = 16 + _ENTRY.size * len(items)
    off = (head + 15) // 16 * 16
    table, payload = [], []
    for k, a in items:
        if a.ndim > 4 or len(k.encode()) >= 96:
            raise ValueError(f"cannot pack tensor {k} with shape {a.shape}")
        dims = list(a.shape) + [1] * (4 - a.ndim)
        table.append(_ENTRY.pack(k.encode(), a.ndim, *dims, 0, off))
        raw = a.tobytes()
        pad = (-len(raw)) % 16
        payload.append(raw + b"\0" * pad)
        off += len(raw) + pad
    blob = struct.pack("<4sIII", b"B2SW", 1, len(items), 0) + b"".join(table)
    blob += b"\0" * ((-len(blob)) % 16)
    return blob + b"".join(payload)
