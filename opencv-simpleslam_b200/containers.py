"""Array-native stand-ins for `list[cv2.KeyPoint]` / `list[cv2.DMatch]`.

The reference builds 2N Python objects per frame (`features_utils.py:61-63, 82`) and its consumers only ever read
`kp[i].pt` (`two_view_bootstrap.py:416-417`, `pnp_utils.py:71`, `keyframe_utils.py:79-80`, `ba_utils.py:122,275`) and
`m.queryIdx / m.trainIdx`.  These sequences keep the data in one ndarray, satisfy len / indexing / slicing / iteration
with the same attribute names, and convert to the real OpenCV objects on request (`to_cv()`).  Opt-in:
`args.array_native = True` for `feature_extractor` / `feature_matcher`; the default stays real lists."""
from __future__ import annotations

from collections.abc import Sequence
from itertools import repeat

import cv2
import numpy as np


class _KeyPointView:
    """One element of a KeyPointArray: the fields of cv2.KeyPoint(x, y, 1)."""
    __slots__ = ("pt", "size", "angle", "response", "octave", "class_id")

    def __init__(self, x, y):
        self.pt, self.size, self.angle, self.response, self.octave, self.class_id = (x, y), 1.0, -1.0, 0.0, 0, -1

    def __repr__(self):
        return f"KeyPointView(pt={self.pt})"


class KeyPointArray(Sequence):
    def __init__(self, pts):
        self.pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)

    def __len__(self):
        return len(self.pts)

    def __getitem__(self, i):
        if isinstance(i, slice) or isinstance(i, (list, np.ndarray)):
            return KeyPointArray(self.pts[i])
        x, y = self.pts[i]
        return _KeyPointView(float(x), float(y))

    def __iter__(self):
        return map(_KeyPointView, self.pts[:, 0].tolist(), self.pts[:, 1].tolist())

    def to_cv(self):
        return list(map(cv2.KeyPoint, self.pts[:, 0].tolist(), self.pts[:, 1].tolist(), repeat(1.0)))


class _DMatchView:
    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, q, t):
        self.queryIdx, self.trainIdx, self.imgIdx, self.distance = q, t, 0, 0.0

    def __repr__(self):
        return f"DMatchView({self.queryIdx}, {self.trainIdx})"


class DMatchArray(Sequence):
    def __init__(self, pairs):
        self.pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)

    queryIdx = property(lambda self: self.pairs[:, 0])
    trainIdx = property(lambda self: self.pairs[:, 1])

    def __len__(self):
        return len(self.pairs)

    def __getitem__(self, i):
        if isinstance(i, slice) or isinstance(i, (list, np.ndarray)):
            return DMatchArray(self.pairs[i])
        q, t = self.pairs[i]
        return _DMatchView(int(q), int(t))

    def __iter__(self):
        return map(_DMatchView, self.pairs[:, 0].tolist(), self.pairs[:, 1].tolist())

    def to_cv(self):
        out = list(map(cv2.DMatch, self.pairs[:, 0].tolist(), self.pairs[:, 1].tolist(), repeat(0.0)))
        for m in out:
            m.imgIdx = 0
        return out
