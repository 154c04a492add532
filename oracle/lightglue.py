"""Oracle: LightGlue matcher on CPU (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates upstream `lightglue/lightglue.py` (un-vendored third-party, called from
`/root/reference/slam/core/features_utils.py:26,157-161`) on its non-compiled,
non-masked, CPU branch, per SURVEY.md Appendix A.3.  Module/parameter names mirror
the upstream state-dict layout.  Independent cross-check:
tests/test_oracle_pins.py runs the HuggingFace `transformers` LightGlue port on the
same weights.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def normalize_keypoints(kpts, size=None):
    if size is None:
        size = 1 + kpts.max(-2).values - kpts.min(-2).values
    elif not isinstance(size, torch.Tensor):
        size = torch.tensor(size, dtype=kpts.dtype)
    size = size.to(kpts)
    shift = size / 2
    scale = size.max(-1).values / 2
    return (kpts - shift[..., None, :]) / scale[..., None, None]


def rotate_half(x):
    x = x.unflatten(-1, (-1, 2))
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).flatten(start_dim=-2)


def apply_cached_rotary_emb(freqs, t):
    return (t * freqs[0]) + (rotate_half(t) * freqs[1])


class LearnableFourierPositionalEncoding(nn.Module):
    def __init__(self, M, dim):
        super().__init__()
        self.Wr = nn.Linear(M, dim // 2, bias=False)

    def forward(self, x):
        projected = self.Wr(x)
        emb = torch.stack([torch.cos(projected), torch.sin(projected)], 0).unsqueeze(-3)
        return emb.repeat_interleave(2, dim=-1)


class TokenConfidence(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.token = nn.Sequential(nn.Linear(dim, 1), nn.Sigmoid())

    def forward(self, desc0, desc1):
        return self.token(desc0).squeeze(-1), self.token(desc1).squeeze(-1)


def _ffn(dim):
    return nn.Sequential(nn.Linear(2 * dim, 2 * dim), nn.LayerNorm(2 * dim, elementwise_affine=True),
                         nn.GELU(), nn.Linear(2 * dim, dim))


class SelfBlock(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.Wqkv = nn.Linear(dim, 3 * dim)
        self.out_proj = nn.Linear(dim, dim)
        self.ffn = _ffn(dim)

    def forward(self, x, encoding, taps=None):
        qkv = self.Wqkv(x)
        qkv = qkv.unflatten(-1, (self.heads, -1, 3)).transpose(1, 2)
        q, k, v = qkv[..., 0], qkv[..., 1], qkv[..., 2]
        q = apply_cached_rotary_emb(encoding, q)
        k = apply_cached_rotary_emb(encoding, k)
        ctx = F.scaled_dot_product_attention(q.contiguous(), k.contiguous(), v.contiguous())
        msg = self.out_proj(ctx.transpose(1, 2).flatten(start_dim=-2))
        if taps is not None:
            taps.update(q=q, k=k, v=v, ctx=ctx, msg=msg)
        return x + self.ffn(torch.cat([x, msg], -1))


class CrossBlock(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.scale = (dim // heads) ** -0.5
        self.to_qk = nn.Linear(dim, dim)
        self.to_v = nn.Linear(dim, dim)
        self.to_out = nn.Linear(dim, dim)
        self.ffn = _ffn(dim)

    def forward(self, x0, x1):
        sp = lambda t: t.unflatten(-1, (self.heads, -1)).transpose(1, 2)  # noqa: E731
        qk0, qk1, v0, v1 = sp(self.to_qk(x0)), sp(self.to_qk(x1)), sp(self.to_v(x0)), sp(self.to_v(x1))
        qk0, qk1 = qk0 * self.scale ** 0.5, qk1 * self.scale ** 0.5
        sim = torch.einsum("bhid, bhjd -> bhij", qk0, qk1)
        attn01 = F.softmax(sim, dim=-1)
        attn10 = F.softmax(sim.transpose(-2, -1).contiguous(), dim=-1)
        m0 = torch.einsum("bhij, bhjd -> bhid", attn01, v1)
        m1 = torch.einsum("bhji, bhjd -> bhid", attn10.transpose(-2, -1), v0)
        mg = lambda t: t.transpose(1, 2).flatten(start_dim=-2)  # noqa: E731
        m0, m1 = self.to_out(mg(m0)), self.to_out(mg(m1))
        x0 = x0 + self.ffn(torch.cat([x0, m0], -1))
        x1 = x1 + self.ffn(torch.cat([x1, m1], -1))
        return x0, x1


class TransformerLayer(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.self_attn = SelfBlock(dim, heads)
        self.cross_attn = CrossBlock(dim, heads)

    def forward(self, desc0, desc1, enc0, enc1):
        desc0 = self.self_attn(desc0, enc0)
        desc1 = self.self_attn(desc1, enc1)
        return self.cross_attn(desc0, desc1)


def sigmoid_log_double_softmax(sim, z0, z1):
    b, m, n = sim.shape
    certainties = F.logsigmoid(z0) + F.logsigmoid(z1).transpose(1, 2)
    scores0 = F.log_softmax(sim, 2)
    scores1 = F.log_softmax(sim.transpose(-1, -2).contiguous(), 2).transpose(-1, -2)
    scores = sim.new_full((b, m + 1, n + 1), 0)
    scores[:, :m, :n] = scores0 + scores1 + certainties
    scores[:, :-1, -1] = F.logsigmoid(-z0.squeeze(-1))
    scores[:, -1, :-1] = F.logsigmoid(-z1.squeeze(-1))
    return scores


class MatchAssignment(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim
        self.matchability = nn.Linear(dim, 1, bias=True)
        self.final_proj = nn.Linear(dim, dim, bias=True)

    def forward(self, desc0, desc1):
        md0, md1 = self.final_proj(desc0), self.final_proj(desc1)
        d = md0.shape[-1]
        md0, md1 = md0 / d ** 0.25, md1 / d ** 0.25
        sim = torch.einsum("bmd,bnd->bmn", md0, md1)
        z0, z1 = self.matchability(desc0), self.matchability(desc1)
        return sigmoid_log_double_softmax(sim, z0, z1), sim

    def get_matchability(self, desc):
        return torch.sigmoid(self.matchability(desc)).squeeze(-1)


def filter_matches(scores, th):
    max0, max1 = scores[:, :-1, :-1].max(2), scores[:, :-1, :-1].max(1)
    m0, m1 = max0.indices, max1.indices
    indices0 = torch.arange(m0.shape[1])[None]
    indices1 = torch.arange(m1.shape[1])[None]
    mutual0 = indices0 == m1.gather(1, m0)
    mutual1 = indices1 == m0.gather(1, m1)
    max0_exp = max0.values.exp()
    zero = max0_exp.new_tensor(0)
    mscores0 = torch.where(mutual0, max0_exp, zero)
    mscores1 = torch.where(mutual1, mscores0.gather(1, m1), zero)
    valid0 = mutual0 & (mscores0 > th)
    valid1 = mutual1 & valid0.gather(1, m1)
    m0 = torch.where(valid0, m0, -1)
    m1 = torch.where(valid1, m1, -1)
    return m0, m1, mscores0, mscores1


class LightGlue(nn.Module):
    """lightglue.LightGlue(features='aliked') stand-in (`features_utils.py:26`)."""

    pruning_keypoint_thresholds = {"cpu": -1, "mps": -1, "cuda": 1024, "flash": 1536}

    def __init__(self, features="aliked", input_dim=128, descriptor_dim=256, n_layers=9, num_heads=4,
                 depth_confidence=0.95, width_confidence=0.99, filter_threshold=0.1, pruning_threshold=None):
        super().__init__()
        assert features == "aliked"
        self.n_layers, self.depth_confidence = n_layers, depth_confidence
        self.width_confidence, self.filter_threshold = width_confidence, filter_threshold
        self.pruning_threshold = self.pruning_keypoint_thresholds["cpu"] if pruning_threshold is None else pruning_threshold
        d = descriptor_dim
        self.input_proj = nn.Linear(input_dim, d, bias=True) if input_dim != d else nn.Identity()
        self.posenc = LearnableFourierPositionalEncoding(2, d // num_heads)
        self.transformers = nn.ModuleList([TransformerLayer(d, num_heads) for _ in range(n_layers)])
        self.log_assignment = nn.ModuleList([MatchAssignment(d) for _ in range(n_layers)])
        self.token_confidence = nn.ModuleList([TokenConfidence(d) for _ in range(n_layers - 1)])
        self.register_buffer("confidence_thresholds", torch.Tensor(
            [self.confidence_threshold(i) for i in range(n_layers)]))
        self.record_taps = False
        self.taps = {}

    def confidence_threshold(self, layer_index):
        return np.clip(0.8 + 0.1 * np.exp(-4.0 * layer_index / self.n_layers), 0, 1)

    def get_pruning_mask(self, confidences, scores, layer_index):
        keep = scores > (1 - self.width_confidence)
        if confidences is not None:
            keep |= confidences <= self.confidence_thresholds[layer_index]
        return keep

    def check_if_stop(self, c0, c1, layer_index, num_points):
        confidences = torch.cat([c0, c1], -1)
        threshold = self.confidence_thresholds[layer_index]
        ratio_confident = 1.0 - (confidences < threshold).float().sum() / num_points
        return ratio_confident > self.depth_confidence

    @torch.no_grad()
    def forward(self, data):
        data0, data1 = data["image0"], data["image1"]
        kpts0, kpts1 = data0["keypoints"], data1["keypoints"]
        b, m, _ = kpts0.shape
        b, n, _ = kpts1.shape
        size0, size1 = data0.get("image_size"), data1.get("image_size")
        kpts0 = normalize_keypoints(kpts0, size0).clone()
        kpts1 = normalize_keypoints(kpts1, size1).clone()
        desc0 = data0["descriptors"].detach().contiguous()
        desc1 = data1["descriptors"].detach().contiguous()
        desc0, desc1 = self.input_proj(desc0), self.input_proj(desc1)
        enc0, enc1 = self.posenc(kpts0), self.posenc(kpts1)
        t = self.taps if self.record_taps else None
        if t is not None:
            t.update(kn0=kpts0, kn1=kpts1, enc0=enc0, enc1=enc1, proj0=desc0, proj1=desc1, layers=[])
        do_early_stop = self.depth_confidence > 0
        do_point_pruning = self.width_confidence > 0
        pruning_th = self.pruning_threshold
        if do_point_pruning:
            ind0 = torch.arange(0, m)[None]
            ind1 = torch.arange(0, n)[None]
            prune0, prune1 = torch.ones_like(ind0), torch.ones_like(ind1)
        token0, token1 = None, None
        i = 0
        for i in range(self.n_layers):
            if desc0.shape[1] == 0 or desc1.shape[1] == 0:
                break
            desc0, desc1 = self.transformers[i](desc0, desc1, enc0, enc1)
            if t is not None:
                t["layers"].append((desc0.clone(), desc1.clone()))
            if i == self.n_layers - 1:
                continue
            if do_early_stop:
                token0, token1 = self.token_confidence[i](desc0, desc1)
                if self.check_if_stop(token0[..., :m], token1[..., :n], i, m + n):
                    break
            if do_point_pruning and desc0.shape[-2] > pruning_th:
                scores0 = self.log_assignment[i].get_matchability(desc0)
                keep0 = torch.where(self.get_pruning_mask(token0, scores0, i))[1]
                ind0 = ind0.index_select(1, keep0)
                desc0 = desc0.index_select(1, keep0)
                enc0 = enc0.index_select(-2, keep0)
                prune0[:, ind0] += 1
            if do_point_pruning and desc1.shape[-2] > pruning_th:
                scores1 = self.log_assignment[i].get_matchability(desc1)
                keep1 = torch.where(self.get_pruning_mask(token1, scores1, i))[1]
                ind1 = ind1.index_select(1, keep1)
                desc1 = desc1.index_select(1, keep1)
                enc1 = enc1.index_select(-2, keep1)
                prune1[:, ind1] += 1
        if desc0.shape[1] == 0 or desc1.shape[1] == 0:
            m0 = desc0.new_full((b, m), -1, dtype=torch.long)
            m1 = desc0.new_full((b, n), -1, dtype=torch.long)
            ms0, ms1 = desc0.new_zeros((b, m)), desc0.new_zeros((b, n))
            matches = desc0.new_empty((b, 0, 2), dtype=torch.long)
            mscores = desc0.new_empty((b, 0))
            if not do_point_pruning:
                prune0 = torch.ones_like(ms0) * self.n_layers
                prune1 = torch.ones_like(ms1) * self.n_layers
            return {"matches0": m0, "matches1": m1, "matching_scores0": ms0, "matching_scores1": ms1,
                    "stop": i + 1, "matches": matches, "scores": mscores, "prune0": prune0, "prune1": prune1}
        desc0, desc1 = desc0[..., :m, :], desc1[..., :n, :]
        scores, sim = self.log_assignment[i](desc0, desc1)
        if t is not None:
            t.update(log_assignment=scores, sim=sim)
        m0, m1, mscores0, mscores1 = filter_matches(scores, self.filter_threshold)
        matches, mscores = [], []
        for k in range(b):
            valid = m0[k] > -1
            mi0 = torch.where(valid)[0]
            mi1 = m0[k][valid]
            if do_point_pruning:
                mi0 = ind0[k, mi0]
                mi1 = ind1[k, mi1]
            matches.append(torch.stack([mi0, mi1], -1))
            mscores.append(mscores0[k][valid])
        if do_point_pruning:
            m0_ = torch.full((b, m), -1, dtype=m0.dtype)
            m1_ = torch.full((b, n), -1, dtype=m1.dtype)
            m0_[:, ind0] = torch.where(m0 == -1, -1, ind1.gather(1, m0.clamp(min=0)))
            m1_[:, ind1] = torch.where(m1 == -1, -1, ind0.gather(1, m1.clamp(min=0)))
            ms0_, ms1_ = torch.zeros((b, m)), torch.zeros((b, n))
            ms0_[:, ind0] = mscores0
            ms1_[:, ind1] = mscores1
            m0, m1, mscores0, mscores1 = m0_, m1_, ms0_, ms1_
        else:
            prune0 = torch.ones_like(mscores0) * self.n_layers
            prune1 = torch.ones_like(mscores1) * self.n_layers
        return {"matches0": m0, "matches1": m1, "matching_scores0": mscores0, "matching_scores1": mscores1,
                "stop": i + 1, "matches": matches, "scores": mscores, "prune0": prune0, "prune1": prune1}
