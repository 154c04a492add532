"""Deterministic synthetic frames for tests and bench: a corner-rich world
canvas and KITTI-shaped crops that overlap so that true correspondences exist.  There is
no network for datasets; BASELINE.json's configs are all stated on synthetic frames."""
from __future__ import annotations

import functools

import cv2
import numpy as np


@functools.lru_cache(maxsize=4)
def world_canvas(height=1024, width=8192, seed=20251017, n_shapes=6000) -> np.ndarray:
    rng = np.random.default_rng(seed)
    low = rng.integers(0, 256, (height // 8, width // 8, 3), dtype=np.uint8)
    img = cv2.resize(low, (width, height), interpolation=cv2.INTER_CUBIC)
    for _ in range(n_shapes):
        kind = int(rng.integers(0, 3))
        col = tuple(int(c) for c in rng.integers(0, 256, 3))
        x, y = int(rng.integers(0, width)), int(rng.integers(0, height))
        s = int(rng.integers(3, 41))
        if kind == 0:
            cv2.rectangle(img, (x, y), (x + s, y + int(rng.integers(3, 41))), col, -1)
        elif kind == 1:
            cv2.circle(img, (x, y), s // 2 + 1, col, -1)
        else:
            a = rng.uniform(0, 2 * np.pi)
            cv2.line(img, (x, y), (int(x + s * np.cos(a)), int(y + s * np.sin(a))), col,
                     int(rng.integers(1, 4)))
    noise = rng.normal(0, 2, img.shape)
    return np.clip(img.astype(np.float64) + noise, 0, 255).astype(np.uint8)


def frame(t: int, height=376, width=1241, canvas=None) -> np.ndarray:
    """BGR u8 [height,width,3]; frame t is the canvas cropped at x = 4t (consecutive frames
    overlap by width-4 px), y wobbling by +-3 px."""
    if canvas is None:
        canvas = world_canvas() if height <= 1024 and width <= 8192 else world_canvas(1200, 12288, 20251018)
    ch, cw = canvas.shape[:2]
    x0 = (4 * t) % (cw - width)
    y0 = (ch - height) // 2 + int(round(3 * np.sin(t / 20.0)))
    return np.ascontiguousarray(canvas[y0:y0 + height, x0:x0 + width])
