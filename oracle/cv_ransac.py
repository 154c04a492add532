"""TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of `cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, 0.99)` exactly as the reference calls it
in `filter_matches_ransac` (/root/reference/slam/core/features_utils.py:185-200).

The algorithm lives in a third-party dependency that is not vendored in /root/reference: OpenCV (requirements.txt
`opencv-python`; the copy in this image is opencv-python-headless 4.13.0.92).  What is restated here is OpenCV's published
implementation, modules/calib3d/src:

* fundam.cpp   `findFundamentalMat`: points converted to float32; n < 7 -> no result; n == 7 -> the 7-point models
               themselves; 8 <= n < 15 -> LMedS even when FM_RANSAC is asked for; n >= 15 -> RANSAC, 7-point model,
               maxIters 1000; `FMEstimatorCallback` (`checkSubset` = `haveCollinearPoints` on both point sets,
               `run7Point`, `computeError` = max of the two squared point-to-epipolar-line distances as float32);
* ptsetreg.cpp `RANSACPointSetRegistrator::run` (RNG seeded with (uint64)-1 on every call, `getSubset`, best model =
               first one with strictly more inliers, `RANSACUpdateNumIters` after every improvement),
               `LMeDSPointSetRegistrator::run`;
* core         `RNG` (multiply-with-carry, 4164903690), `solveCubic`, and the part of `SVD::compute(FULL_UV)` that
               matters here: the two null-space rows of Vt of a 7 x 9 matrix are the fixed +-1/9 sign vectors drawn
               from `RNG(0x12345678)`, projected off the row space and orthonormalised (JacobiSVDImpl_'s completion
               loop) - so they are a function of the row space only and can be computed without Jacobi rotations.

PARITY IS PINNED: tests/test_oracle_cv_ransac.py runs cv2.findFundamentalMat itself (cv2 IS the reference's
implementation of this row) on seeded scenes and requires identical masks and F to 1e-9; the null-space basis is
compared with cv2.SVDecomp directly.  The committed fixtures tests/golden/fm_cv_*.npz were produced by cv2
(tests/golden/make_golden_fm_cv.py).
"""
from __future__ import annotations

import math

import numpy as np

M32 = 0xFFFFFFFF
M64 = (1 << 64) - 1
FLT_EPSILON = float(np.finfo(np.float32).eps)
DBL_EPSILON = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)
MODEL_POINTS = 7
MAX_ITERS = 1000           # findFundamentalMat's default maxIters (the reference does not pass one)
SUBSET_ATTEMPTS = 10000    # RANSAC run() -> getSubset(..., 10000); LMedS uses getSubset's default 1000


class CvRNG:
    """cv::RNG: state = (uint32)state * 4164903690 + (state >> 32); next() returns the low 32 bits."""

    def __init__(self, state: int):
        self.state = (state & M64) or 0xFFFFFFFF

    def next(self) -> int:
        self.state = ((self.state & M32) * 4164903690 + (self.state >> 32)) & M64
        return self.state & M32

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else a + self.next() % (b - a)


def null_space_signs():
    """The two +-1/9 vectors JacobiSVDImpl_ draws (RNG(0x12345678), `(rng.next() & 256) != 0 ? val0 : -val0`) for rows 7
    and 8 of Vt of a 7 x 9 matrix of full row rank."""
    rng = CvRNG(0x12345678)
    r = [[(1.0 / 9 if rng.next() & 256 else -1.0 / 9) for _ in range(9)] for _ in range(2)]
    return np.array(r[0]), np.array(r[1])


_R1, _R2 = null_space_signs()


def null_space_basis(A: np.ndarray):
    """Rows 7 and 8 of Vt from SVDecomp(A[7,9], MODIFY_A + FULL_UV): orthonormalise the rows of A (modified Gram-Schmidt,
    two passes), then f1 = unit(P r1), f2 = unit(P r2 - (f1 . P r2) f1) with P the projector onto the null space."""
    Q = []
    for i in range(7):
        v = A[i].astype(np.float64).copy()
        for _ in range(2):
            for q in Q:
                v = v - (q @ v) * q
        nv = math.sqrt(float(v @ v))
        if nv > 0:
            Q.append(v / nv)

    def project(r, extra=()):
        v = r.copy()
        for _ in range(2):
            for q in list(Q) + list(extra):
                v = v - (q @ v) * q
        return v / math.sqrt(float(v @ v))

    f1 = project(_R1)
    f2 = project(_R2, (f1,))
    return f1, f2


def solve_cubic(c):
    """cv::solveCubic(coeffs[4]) for c0 x^3 + c1 x^2 + c2 x + c3: (n, roots) in OpenCV's root order."""
    a0, a1, a2, a3 = (float(v) for v in c)
    x = [0.0, 0.0, 0.0]
    n = 0
    if a0 == 0:
        if a1 == 0:
            if a2 == 0:
                n = -1 if a3 == 0 else 0
            else:
                x[0] = -a3 / a2
                n = 1
        else:
            d = a2 * a2 - 4 * a1 * a3
            if d >= 0:
                d = math.sqrt(d)
                q1 = (-a2 + d) * 0.5
                q2 = (a2 + d) * -0.5
                if abs(q1) > abs(q2):
                    x[0] = q1 / a1
                    x[1] = a3 / q1
                else:
                    x[0] = q2 / a1
                    x[1] = a3 / q2
                n = 2 if d > 0 else 1
    else:
        a0 = 1.0 / a0
        a1 *= a0
        a2 *= a0
        a3 *= a0
        Q = (a1 * a1 - 3 * a2) * (1.0 / 9)
        R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1.0 / 54)
        Qcubed = Q * Q * Q
        d = (a1 * a1 * (a2 * a2 - 4 * a1 * a3) + 2 * a2 * (9 * a1 * a3 - 2 * a2 * a2) - 27 * a3 * a3) * (1.0 / 108)
        if d > 0:
            theta = math.acos(R / math.sqrt(Qcubed))
            sqrtQ = math.sqrt(Q)
            t0 = -2 * sqrtQ
            t1 = theta * (1.0 / 3)
            t2 = a1 * (1.0 / 3)
            x[0] = t0 * math.cos(t1) - t2
            x[1] = t0 * math.cos(t1 + (2.0 * math.pi / 3)) - t2
            x[2] = t0 * math.cos(t1 + (4.0 * math.pi / 3)) - t2
            n = 3
        elif d == 0:
            if R >= 0:
                x[0] = -2 * R ** (1.0 / 3) - a1 / 3
                x[1] = R ** (1.0 / 3) - a1 / 3
            else:
                x[0] = 2 * (-R) ** (1.0 / 3) - a1 / 3
                x[1] = -((-R) ** (1.0 / 3)) - a1 / 3
            n = 1 if x[0] == x[1] else 2
            x[1] = 0.0 if x[0] == x[1] else x[1]
        else:
            d = math.sqrt(-d)
            e = (d + abs(R)) ** (1.0 / 3)
            if R > 0:
                e = -e
            x[0] = (e + Q / e) - a1 * (1.0 / 3)
            n = 1
    return n, x


def run_7point(m1: np.ndarray, m2: np.ndarray) -> list:
    """fundam.cpp run7Point on seven float32 correspondences: list of 0..3 F (3x3 float64), F[2,2] = 1."""
    m1 = m1.astype(np.float32).astype(np.float64)
    m2 = m2.astype(np.float32).astype(np.float64)
    t = 1.0 / 7
    c1 = np.array([0.0, 0.0])
    c2 = np.array([0.0, 0.0])
    for i in range(7):
        c1 = c1 + m1[i]
        c2 = c2 + m2[i]
    c1 = c1 * t
    c2 = c2 * t
    s1 = s2 = 0.0
    for i in range(7):
        s1 += math.sqrt((m1[i, 0] - c1[0]) ** 2 + (m1[i, 1] - c1[1]) ** 2)
        s2 += math.sqrt((m2[i, 0] - c2[0]) ** 2 + (m2[i, 1] - c2[1]) ** 2)
    s1 *= t
    s2 *= t
    if s1 < FLT_EPSILON or s2 < FLT_EPSILON:
        return []
    s1 = math.sqrt(2.0) / s1
    s2 = math.sqrt(2.0) / s2
    A = np.empty((7, 9))
    for i in range(7):
        x0, y0 = (m1[i, 0] - c1[0]) * s1, (m1[i, 1] - c1[1]) * s1
        x1, y1 = (m2[i, 0] - c2[0]) * s2, (m2[i, 1] - c2[1]) * s2
        A[i] = [x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1.0]
    f1, f2 = null_space_basis(A)
    f1 = f1 - f2
    c = [0.0] * 4
    t0 = f2[4] * f2[8] - f2[5] * f2[7]
    t1 = f2[3] * f2[8] - f2[5] * f2[6]
    t2 = f2[3] * f2[7] - f2[4] * f2[6]
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2
    c[2] = (f1[0] * t0 - f1[1] * t1 + f1[2] * t2
            - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6])
            - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4])
            - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) + f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]))
    t0 = f1[4] * f1[8] - f1[5] * f1[7]
    t1 = f1[3] * f1[8] - f1[5] * f1[6]
    t2 = f1[3] * f1[7] - f1[4] * f1[6]
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2
    c[1] = (f2[0] * t0 - f2[1] * t1 + f2[2] * t2
            - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6])
            - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4])
            - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) + f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]))
    n, r = solve_cubic(c)
    if n < 1 or n > 3:
        return []
    T1 = np.array([[s1, 0, -s1 * c1[0]], [0, s1, -s1 * c1[1]], [0, 0, 1.0]])
    T2 = np.array([[s2, 0, -s2 * c2[0]], [0, s2, -s2 * c2[1]], [0, 0, 1.0]])
    out = []
    for k in range(n):
        lam, mu = r[k], 1.0
        s = f1[8] * r[k] + f2[8]
        fm = np.empty(9)
        if abs(s) > DBL_EPSILON:
            mu = 1.0 / s
            lam *= mu
            fm[8] = 1.0
        else:
            fm[8] = 0.0
        fm[:8] = f1[:8] * lam + f2[:8] * mu
        F = T2.T @ fm.reshape(3, 3) @ T1
        if abs(F[2, 2]) > FLT_EPSILON:
            F = F * (1.0 / F[2, 2])
        out.append(F)
    return out


def compute_error(F: np.ndarray, m1: np.ndarray, m2: np.ndarray) -> np.ndarray:
    """FMEstimatorCallback::computeError: float32 [n]; products and sums in the source's order, no contraction."""
    f = F.reshape(9)
    x1, y1 = m1[:, 0].astype(np.float64), m1[:, 1].astype(np.float64)
    x2, y2 = m2[:, 0].astype(np.float64), m2[:, 1].astype(np.float64)
    a = f[0] * x1 + f[1] * y1 + f[2]
    b = f[3] * x1 + f[4] * y1 + f[5]
    c = f[6] * x1 + f[7] * y1 + f[8]
    with np.errstate(divide="ignore", invalid="ignore"):
        s2 = 1.0 / (a * a + b * b)
        d2 = x2 * a + y2 * b + c
        a = f[0] * x2 + f[3] * y2 + f[6]
        b = f[1] * x2 + f[4] * y2 + f[7]
        c = f[2] * x2 + f[5] * y2 + f[8]
        s1 = 1.0 / (a * a + b * b)
        d1 = x1 * a + y1 * b + c
        return np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(np.float32)


def find_inliers(F, m1, m2, thresh: float):
    t = np.float32(thresh * thresh)
    mask = compute_error(F, m1, m2) <= t      # NaN compares false, like the C++ `errptr[i] <= t`
    return int(mask.sum()), mask.astype(np.uint8)


def have_collinear_points(ms: np.ndarray, count: int) -> bool:
    """fundam.cpp haveCollinearPoints: only the LAST point is tested against the lines through earlier pairs; the
    coordinate differences are float32 subtractions widened to double."""
    i = count - 1
    ms = ms.astype(np.float32)
    for j in range(i):
        dx1 = float(np.float32(ms[j, 0] - ms[i, 0]))
        dy1 = float(np.float32(ms[j, 1] - ms[i, 1]))
        for k in range(j):
            dx2 = float(np.float32(ms[k, 0] - ms[i, 0]))
            dy2 = float(np.float32(ms[k, 1] - ms[i, 1]))
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def get_subset(m1, m2, rng: CvRNG, max_attempts: int):
    """ptsetreg.cpp getSubset: seven distinct uniform draws (redraw on a duplicate), whole subset redrawn when
    checkSubset rejects it.  Returns the index list or None."""
    count = len(m1)
    for _ in range(max_attempts):
        idx = []
        for _i in range(MODEL_POINTS):
            c = rng.uniform(0, count)
            while c in idx:
                c = rng.uniform(0, count)
            idx.append(c)
        if not have_collinear_points(m1[idx], MODEL_POINTS) and not have_collinear_points(m2[idx], MODEL_POINTS):
            return idx
    return None


def ransac_update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < DBL_MIN:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))           # cvRound: round half to even


def subsets(m1, m2, n_iters: int = MAX_ITERS, max_attempts: int = SUBSET_ATTEMPTS):
    """The sample of every iteration up to n_iters.  The RNG stream does not depend on the models, so all samples of a
    call can be drawn before any model is evaluated (what the CUDA path does).  A failed getSubset ends the list."""
    rng = CvRNG(M64)
    out = []
    for _ in range(n_iters):
        idx = get_subset(m1, m2, rng, max_attempts)
        if idx is None:
            break
        out.append(idx)
    return out


def scan_counts(counts, count: int, confidence: float = 0.99, max_iters: int = MAX_ITERS):
    """The sequential part of RANSACPointSetRegistrator::run given every model's inlier count (counts[iter] = list of
    the iteration's model counts, in model order): returns (iter, model, maxGoodCount, niters) of the winner or None."""
    niters = max(max_iters, 1)
    best = None
    max_good = 0
    it = 0
    while it < niters and it < len(counts):
        for k, good in enumerate(counts[it]):
            if good > max(max_good, MODEL_POINTS - 1):
                max_good = good
                best = (it, k)
                niters = ransac_update_num_iters(confidence, (count - good) / count, MODEL_POINTS, niters)
        it += 1
    if best is None:
        return None
    return best[0], best[1], max_good, niters


def _prep(pts1, pts2):
    m1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    m2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    assert len(m1) == len(m2)
    return m1, m2


def ransac_run(m1, m2, thresh: float, confidence: float = 0.99, max_iters: int = MAX_ITERS):
    """RANSACPointSetRegistrator::run.  Returns (F or None, mask uint8 [n] or None, info dict)."""
    count = len(m1)
    niters = max(max_iters, 1)
    rng = CvRNG(M64)
    max_good, best_F, best_mask, best_at = 0, None, None, None
    it = 0
    while it < niters:
        idx = get_subset(m1, m2, rng, SUBSET_ATTEMPTS)
        if idx is None:
            if it == 0:
                return None, None, {"iters": 0}
            break
        for k, F in enumerate(run_7point(m1[idx], m2[idx])):
            good, mask = find_inliers(F, m1, m2, thresh)
            if good > max(max_good, MODEL_POINTS - 1):
                max_good, best_F, best_mask, best_at = good, F, mask, (it, k)
                niters = ransac_update_num_iters(confidence, (count - good) / count, MODEL_POINTS, niters)
        it += 1
    if max_good <= 0:
        return None, None, {"iters": it}
    return best_F, best_mask, {"iters": it, "winner": best_at, "count": max_good, "niters": niters}


def lmeds_run(m1, m2, confidence: float = 0.99, max_iters: int = MAX_ITERS):
    """LMeDSPointSetRegistrator::run (what FM_RANSAC silently becomes for 8 <= n < 15)."""
    count = len(m1)
    rng = CvRNG(M64)
    niters = max(ransac_update_num_iters(confidence, 0.45, MODEL_POINTS, max_iters), 3)
    min_median, best_F, best_at = float("inf"), None, None
    for it in range(niters):
        idx = get_subset(m1, m2, rng, 1000)
        if idx is None:
            if it == 0:
                return None, None, {"iters": 0}
            break
        for k, F in enumerate(run_7point(m1[idx], m2[idx])):
            err = compute_error(F, m1, m2)
            median = float(np.sort(err.view(np.int32))[count // 2].view(np.float32))   # nth_element on the float bits
            if median < min_median:
                min_median, best_F, best_at = median, F, (it, k)
    if best_F is None:
        return None, None, {"iters": niters}
    sigma = max(2.5 * 1.4826 * (1 + 5.0 / (count - MODEL_POINTS)) * math.sqrt(min_median), 0.001)
    good, mask = find_inliers(best_F, m1, m2, sigma)
    info = {"iters": niters, "winner": best_at, "count": good, "median": min_median, "sigma": sigma}
    if good < MODEL_POINTS:
        return None, mask, info                 # cv2 still hands back the mask it filled
    return best_F, mask, info


def find_fundamental_mat(pts1, pts2, thresh: float = 3.0, confidence: float = 0.99, max_iters: int = MAX_ITERS):
    """cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, confidence) -> (F, mask [n,1] uint8) or (None, None)."""
    m1, m2 = _prep(pts1, pts2)
    n = len(m1)
    if n < 7:
        return None, None
    if n == 7:
        Fs = run_7point(m1, m2)
        if not Fs:
            return None, np.ones((7, 1), np.uint8)
        return np.concatenate(Fs, axis=0), np.ones((7, 1), np.uint8)
    if thresh <= 0:
        thresh = 3.0
    if confidence < DBL_EPSILON or confidence > 1 - DBL_EPSILON:
        confidence = 0.99
    if n >= 15:
        F, mask, _ = ransac_run(m1, m2, thresh, confidence, max_iters)
    else:
        F, mask, _ = lmeds_run(m1, m2, confidence, max_iters)
    if F is None:
        return None, (mask.reshape(-1, 1) if mask is not None else None)
    return F, mask.reshape(-1, 1)
