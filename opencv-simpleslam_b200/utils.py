"""`lightglue.utils` look-alike: the two names the reference imports at
/root/reference/slam/core/features_utils.py:9 (`rbd`, `load_image`; the latter is imported
there but never called)."""
import cv2
import numpy as np
import torch

from .frontend import rbd  # noqa: F401


def load_image(path, resize=None, **_):
    """upstream lightglue.utils.load_image: file -> RGB float tensor [3,H,W] in [0,1]."""
    img = cv2.imread(str(path), cv2.IMREAD_COLOR)
    if img is None:
        raise IOError(f"Could not read image at {path}.")
    img = img[..., ::-1]
    if resize is not None:
        img = cv2.resize(img, (resize, resize) if isinstance(resize, int) else tuple(resize))
    return torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)).astype(np.float32) / 255.0)
