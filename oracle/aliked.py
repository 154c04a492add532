"""Oracle: ALIKED extractor on CPU (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates upstream `lightglue/aliked.py` + `lightglue/utils.py::Extractor.extract`
(un-vendored third-party, called from
`/root/reference/slam/core/features_utils.py:25,94`) per SURVEY.md Appendix A.1/A.2.
The module tree and parameter names mirror the upstream state-dict layout
 so a real `aliked-n16.pth` loads with strict=True.
`stage taps`: every intermediate tensor is stored in `self.taps` when
`self.record_taps` is set, so each CUDA kernel can be parity-tested alone.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.ops import deform_conv2d

from .preprocess import resize_long_side

CFGS = {  # c1, c2, c3, c4, dim, K, M
    "aliked-t16": (8, 16, 32, 64, 64, 3, 16),
    "aliked-n16": (16, 32, 64, 128, 128, 3, 16),
    "aliked-n16rot": (16, 32, 64, 128, 128, 3, 16),
    "aliked-n32": (16, 32, 64, 128, 128, 3, 32),
}


def simple_nms(scores: torch.Tensor, nms_radius: int) -> torch.Tensor:
    """upstream `simple_nms` (== HF SuperPoint `simple_nms`)."""
    def mp(x):
        return F.max_pool2d(x, kernel_size=nms_radius * 2 + 1, stride=1, padding=nms_radius)

    zeros = torch.zeros_like(scores)
    max_mask = scores == mp(scores)
    for _ in range(2):
        supp_mask = mp(max_mask.float()) > 0
        supp_scores = torch.where(supp_mask, zeros, scores)
        new_max_mask = supp_scores == mp(supp_scores)
        max_mask = max_mask | (new_max_mask & (~supp_mask))
    return torch.where(max_mask, scores, zeros)


def get_patches(tensor: torch.Tensor, required_corners: torch.Tensor, ps: int) -> torch.Tensor:
    """upstream `get_patches`: [C,H,W], int corners [N,2](x,y) -> [N,C,ps(dy),ps(dx)]."""
    c, h, w = tensor.shape
    corner = (required_corners - ps / 2 + 1).long()
    corner[:, 0] = corner[:, 0].clamp(min=0, max=w - 1 - ps)
    corner[:, 1] = corner[:, 1].clamp(min=0, max=h - 1 - ps)
    offset = torch.arange(0, ps)
    x, y = torch.meshgrid(offset, offset, indexing="ij")
    patches = torch.stack((x, y)).permute(2, 1, 0).unsqueeze(2)
    patches = patches.to(corner) + corner[None, None]
    pts = patches.reshape(-1, 2)
    sampled = tensor.permute(1, 2, 0)[tuple(pts.T)[::-1]]
    sampled = sampled.reshape(ps, ps, -1, c)
    return sampled.permute(2, 3, 0, 1)


class InputPadder:
    def __init__(self, h, w, divis_by=8):
        pad_ht = (((h // divis_by) + 1) * divis_by - h) % divis_by
        pad_wd = (((w // divis_by) + 1) * divis_by - w) % divis_by
        self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]

    def pad(self, x):
        return F.pad(x, self._pad, mode="replicate")

    def unpad(self, x):
        ht, wd = x.shape[-2:]
        c = [self._pad[2], ht - self._pad[3], self._pad[0], wd - self._pad[1]]
        return x[..., c[0]:c[1], c[2]:c[3]]


class DeformableConv2d(nn.Module):
    def __init__(self, cin, cout, kernel_size=3, padding=1):
        super().__init__()
        self.padding = padding
        self.offset_conv = nn.Conv2d(cin, 2 * kernel_size * kernel_size, kernel_size, 1, padding, bias=True)
        self.regular_conv = nn.Conv2d(cin, cout, kernel_size, 1, padding, bias=False)

    def forward(self, x):
        h, w = x.shape[2:]
        max_offset = max(h, w) / 4.0
        offset = self.offset_conv(x).clamp(-max_offset, max_offset)
        return deform_conv2d(x, offset, self.regular_conv.weight, None, padding=self.padding)


def _conv(cin, cout, conv_type):
    if conv_type == "conv":
        return nn.Conv2d(cin, cout, 3, 1, 1, bias=False)
    return DeformableConv2d(cin, cout, 3, 1)


class ConvBlock(nn.Module):
    def __init__(self, cin, cout, conv_type="conv"):
        super().__init__()
        self.gate = nn.SELU()
        self.conv1 = _conv(cin, cout, conv_type)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = _conv(cout, cout, conv_type)
        self.bn2 = nn.BatchNorm2d(cout)

    def forward(self, x):
        x = self.gate(self.bn1(self.conv1(x)))
        return self.gate(self.bn2(self.conv2(x)))


class ResBlock(nn.Module):
    def __init__(self, cin, cout, conv_type="conv"):
        super().__init__()
        self.gate = nn.SELU()
        self.conv1 = _conv(cin, cout, conv_type)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = _conv(cout, cout, conv_type)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        out = self.gate(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        out = out + self.downsample(x)
        return self.gate(out)


class SDDH(nn.Module):
    def __init__(self, dims, kernel_size=3, n_pos=8):
        super().__init__()
        self.kernel_size, self.n_pos = kernel_size, n_pos
        ch = 2 * n_pos
        self.offset_conv = nn.Sequential(
            nn.Conv2d(dims, ch, kernel_size, 1, 0, bias=True), nn.SELU(), nn.Conv2d(ch, ch, 1, 1, 0, bias=True))
        self.sf_conv = nn.Conv2d(dims, dims, 1, 1, 0, bias=False)
        self.agg_weights = nn.Parameter(torch.rand(n_pos, dims, dims))

    def forward(self, x, keypoints, taps=None):
        b, c, h, w = x.shape
        wh = torch.tensor([[w - 1, h - 1]], dtype=x.dtype)
        max_offset = max(h, w) / 4.0
        descriptors = []
        for ib in range(b):
            xi, kptsi = x[ib], keypoints[ib]
            kptsi_wh = (kptsi / 2 + 0.5) * wh
            n = len(kptsi)
            patch = get_patches(xi, kptsi_wh.long(), self.kernel_size)
            offset = self.offset_conv(patch).clamp(-max_offset, max_offset)
            offset = offset.view(n, self.n_pos, 2)
            pos = kptsi_wh.unsqueeze(1) + offset
            pos = 2.0 * pos / wh[None] - 1
            pos = pos.reshape(1, n * self.n_pos, 1, 2)
            feats = F.grid_sample(xi.unsqueeze(0), pos, mode="bilinear", align_corners=True)
            feats = feats.reshape(c, n, self.n_pos, 1).permute(1, 0, 2, 3)
            feats = F.selu(self.sf_conv(feats)).squeeze(-1)
            descs = torch.einsum("ncp,pcd->nd", feats, self.agg_weights)
            if taps is not None:
                taps["sddh_offset"] = offset
                taps["sddh_desc_raw"] = descs
            descriptors.append(F.normalize(descs, p=2.0, dim=1))
        return descriptors


class DKD(nn.Module):
    def __init__(self, radius=2, top_k=-1, scores_th=0.2, n_limit=20000):
        super().__init__()
        self.radius, self.top_k, self.scores_th, self.n_limit = radius, top_k, scores_th, n_limit
        self.kernel_size = 2 * radius + 1
        self.temperature = 0.1
        self.unfold = nn.Unfold(kernel_size=self.kernel_size, padding=radius)
        x = torch.linspace(-radius, radius, self.kernel_size)
        self.hw_grid = torch.stack(torch.meshgrid([x, x], indexing="ij")).view(2, -1).t()[:, [1, 0]]

    def forward(self, scores_map, taps=None):
        b, c, h, w = scores_map.shape
        nms_scores = simple_nms(scores_map, self.radius)
        r = self.radius
        nms_scores[:, :, :r, :] = 0
        nms_scores[:, :, :, :r] = 0
        nms_scores[:, :, -r:, :] = 0
        nms_scores[:, :, :, -r:] = 0
        if self.top_k > 0:
            topk = torch.topk(nms_scores.view(b, -1), self.top_k)
            indices_keypoints = [topk.indices[i] for i in range(b)]
        else:
            if self.scores_th > 0:
                masks = nms_scores > self.scores_th
                if masks.sum() == 0:
                    th = scores_map.reshape(b, -1).mean(dim=1)
                    masks = nms_scores > th.reshape(b, 1, 1, 1)
            else:
                th = scores_map.reshape(b, -1).mean(dim=1)
                masks = nms_scores > th.reshape(b, 1, 1, 1)
            masks = masks.reshape(b, -1)
            indices_keypoints = []
            for mask, scores in zip(masks, scores_map.reshape(b, -1)):
                indices = mask.nonzero()[:, 0]
                if taps is not None:
                    taps["dkd_candidates"] = indices
                if len(indices) > self.n_limit:
                    # stable=True pins the tie-break (score desc, then raster index asc);
                    # upstream's sort is unspecified on ties.
                    sort_idx = scores[indices].sort(descending=True, stable=True)[1]
                    indices = indices[sort_idx[: self.n_limit]]
                indices_keypoints.append(indices)
        wh = torch.tensor([w - 1, h - 1], dtype=scores_map.dtype)
        keypoints, dispersitys, kptscores = [], [], []
        patches = self.unfold(scores_map)
        grid = self.hw_grid.to(scores_map)
        for bi in range(b):
            patch = patches[bi].t()
            idx = indices_keypoints[bi]
            patch_scores = patch[idx]
            xy_nms = torch.stack([idx % w, torch.div(idx, w, rounding_mode="trunc")], dim=1)
            max_v = patch_scores.max(dim=1).values[:, None]
            x_exp = ((patch_scores - max_v) / self.temperature).exp()
            xy_res = x_exp @ grid / x_exp.sum(dim=1)[:, None]
            dist2 = torch.norm((grid[None] - xy_res[:, None]) / self.radius, dim=-1) ** 2
            disp = (x_exp * dist2).sum(dim=1) / x_exp.sum(dim=1)
            xy = xy_nms + xy_res
            xy = xy / wh * 2 - 1
            kptscore = F.grid_sample(scores_map[bi].unsqueeze(0), xy.view(1, 1, -1, 2),
                                     mode="bilinear", align_corners=True)[0, 0, 0, :]
            keypoints.append(xy)
            dispersitys.append(disp)
            kptscores.append(kptscore)
            if taps is not None:
                taps["dkd_indices"] = idx
        return keypoints, dispersitys, kptscores


class ALIKED(nn.Module):
    """lightglue.ALIKED stand-in; same constructor keywords that the reference uses
    (`features_utils.py:25`: `ALIKED(max_num_keypoints=...)`)."""

    def __init__(self, model_name="aliked-n16", max_num_keypoints=-1, detection_threshold=0.2,
                 nms_radius=2, resize=1024):
        super().__init__()
        c1, c2, c3, c4, dim, K, M = CFGS[model_name]
        self.resize = resize
        self.gate = nn.SELU()
        self.pool2 = nn.AvgPool2d(2, 2)
        self.pool4 = nn.AvgPool2d(4, 4)
        self.block1 = ConvBlock(3, c1, "conv")
        self.block2 = ResBlock(c1, c2, "conv")
        self.block3 = ResBlock(c2, c3, "dcn")
        self.block4 = ResBlock(c3, c4, "dcn")
        self.conv1 = nn.Conv2d(c1, dim // 4, 1, bias=False)
        self.conv2 = nn.Conv2d(c2, dim // 4, 1, bias=False)
        self.conv3 = nn.Conv2d(c3, dim // 4, 1, bias=False)
        self.conv4 = nn.Conv2d(c4, dim // 4, 1, bias=False)
        self.upsample2 = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.upsample8 = nn.Upsample(scale_factor=8, mode="bilinear", align_corners=True)
        self.upsample32 = nn.Upsample(scale_factor=32, mode="bilinear", align_corners=True)
        self.score_head = nn.Sequential(
            nn.Conv2d(dim, 8, 1, bias=False), nn.SELU(),
            nn.Conv2d(8, 4, 3, 1, 1, bias=False), nn.SELU(),
            nn.Conv2d(4, 4, 3, 1, 1, bias=False), nn.SELU(),
            nn.Conv2d(4, 1, 3, 1, 1, bias=False))
        self.desc_head = SDDH(dim, K, M)
        self.dkd = DKD(radius=nms_radius,
                       top_k=-1 if detection_threshold > 0 else max_num_keypoints,
                       scores_th=detection_threshold,
                       n_limit=max_num_keypoints if max_num_keypoints > 0 else 20000)
        self.record_taps = False
        self.taps = {}

    def extract_dense_map(self, image):
        t = self.taps if self.record_taps else None
        padder = InputPadder(image.shape[-2], image.shape[-1], 32)
        image = padder.pad(image)
        x1 = self.block1(image)
        x2 = self.block2(self.pool2(x1))
        x3 = self.block3(self.pool4(x2))
        x4 = self.block4(self.pool4(x3))
        if t is not None:
            t.update(padded=image, x1=x1, x2=x2, x3=x3, x4=x4)
        x1 = self.gate(self.conv1(x1))
        x2 = self.gate(self.conv2(x2))
        x3 = self.gate(self.conv3(x3))
        x4 = self.gate(self.conv4(x4))
        x1234 = torch.cat([x1, self.upsample2(x2), self.upsample8(x3), self.upsample32(x4)], dim=1)
        score_map = torch.sigmoid(self.score_head(x1234))
        feature_map = F.normalize(x1234, p=2, dim=1)
        feature_map = padder.unpad(feature_map)
        score_map = padder.unpad(score_map)
        if t is not None:
            t.update(score_map=score_map, feature_map=feature_map)
        return feature_map, score_map

    def forward(self, data):
        image = data["image"]
        t = self.taps if self.record_taps else None
        feature_map, score_map = self.extract_dense_map(image)
        # upstream DKD returns (keypoints, dispersity, scores) and ALIKED.forward unpacks
        # them as (keypoints, kptscores, scoredispersitys): "keypoint_scores" therefore
        # carries the dispersity ([U?]; the reference never reads it).
        keypoints, kptscores, scoredispersitys = self.dkd(score_map, taps=t)
        descriptors = self.desc_head(feature_map, keypoints, taps=t)
        _, _, h, w = image.shape
        wh = torch.tensor([w - 1, h - 1], dtype=image.dtype)
        if t is not None:
            t.update(kp_norm=keypoints[0], kp_sampled_score=scoredispersitys[0])
        return {
            "keypoints": wh * (torch.stack(keypoints) + 1) / 2.0,
            "descriptors": torch.stack(descriptors),
            "keypoint_scores": torch.stack(kptscores),
        }

    @torch.no_grad()
    def extract(self, img, **conf):
        """upstream `Extractor.extract`: resize -> forward -> keypoints back to original px."""
        if img.dim() == 3:
            img = img[None]
        assert img.dim() == 4 and img.shape[0] == 1
        shape = img.shape[-2:][::-1]
        img, scales = resize_long_side(img, self.resize)
        if self.record_taps:
            self.taps["resized"] = img
        feats = self.forward({"image": img})
        feats["image_size"] = torch.tensor(shape)[None].to(img).float()
        feats["keypoints"] = (feats["keypoints"] + 0.5) / scales[None] - 0.5
        return feats
