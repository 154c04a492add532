"""Oracle: image preprocessing (TEST INFRASTRUCTURE - see oracle/__init__.py).

Follows
  * `/root/reference/slam/core/features_utils.py:219-222`  (`_bgr_to_tensor`)
  * upstream `lightglue/utils.py::ImagePreprocessor` (resize=1024, side="long",
    bilinear, align_corners=None, antialias=True)   [un-vendored, SURVEY A.1]
  * upstream `kornia/geometry/transform/affwarp.py::resize` and
    `kornia/filters/gaussian.py::gaussian_blur2d`   [un-vendored, SURVEY A.1]
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def bgr_to_tensor(image: np.ndarray) -> torch.Tensor:
    """features_utils.py:219-222 without cv2: BGR u8 HxWx3 -> RGB f32 1x3xHxW in [0,1]."""
    rgb = image[..., ::-1].astype(np.float32) / 255.0
    return torch.from_numpy(np.ascontiguousarray(rgb)).permute(2, 0, 1).unsqueeze(0)


def resized_shape(h: int, w: int, size: int = 1024):
    """kornia `_side_to_image_size(size, w/h, side='long')` (int() truncation)."""
    ar = w / h
    if ar >= 1.0:
        return int(size / ar), size
    return size, int(size * ar)


def gaussian_kernel1d(ks: int, sigma: float) -> torch.Tensor:
    x = torch.arange(ks, dtype=torch.float32) - ks // 2
    if ks % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma * sigma))
    return g / g.sum()


def blur_params(h: int, w: int, hn: int, wn: int):
    """(ks_y, ks_x, sigma_y, sigma_x) or None when kornia would not blur."""
    fy, fx = h / hn, w / wn
    if max(fy, fx) <= 1:
        return None
    sy = max((fy - 1.0) / 2.0, 0.001)
    sx = max((fx - 1.0) / 2.0, 0.001)
    ky = int(max(2.0 * 2 * sy, 3))
    kx = int(max(2.0 * 2 * sx, 3))
    if ky % 2 == 0:
        ky += 1
    if kx % 2 == 0:
        kx += 1
    return ky, kx, sy, sx


def gaussian_blur2d(img: torch.Tensor, ky: int, kx: int, sy: float, sx: float) -> torch.Tensor:
    """Separable, reflect border; x pass then y pass (kornia filter2d_separable)."""
    c = img.shape[1]
    gx = gaussian_kernel1d(kx, sx).to(img)
    gy = gaussian_kernel1d(ky, sy).to(img)
    t = F.pad(img, (kx // 2, kx // 2, 0, 0), mode="reflect")
    t = F.conv2d(t, gx.view(1, 1, 1, kx).expand(c, 1, 1, kx), groups=c)
    t = F.pad(t, (0, 0, ky // 2, ky // 2), mode="reflect")
    t = F.conv2d(t, gy.view(1, 1, ky, 1).expand(c, 1, ky, 1), groups=c)
    return t


def resize_long_side(img: torch.Tensor, size: int = 1024):
    """Returns (resized [1,3,H',W'], scales f32 [2] = (W'/W, H'/H))."""
    h, w = img.shape[-2:]
    hn, wn = resized_shape(h, w, size)
    if (hn, wn) != (h, w):
        bp = blur_params(h, w, hn, wn)
        if bp is not None:
            img = gaussian_blur2d(img, *bp)
        img = F.interpolate(img, size=(hn, wn), mode="bilinear", align_corners=None)
    scales = torch.tensor([img.shape[-1] / w, img.shape[-2] / h], dtype=torch.float32)
    return img, scales
