"""Shared test inputs (seeded, regenerated - never read from /root/reference)."""
import os

import numpy as np
import torch


def noisy_copy_pair(m, n, seed=0, noise=0.05):
    """Unit descriptors d0 [m,128]; d1 = normalised noisy copies of a permuted subset (true
    correspondences exist), keypoints in a KITTI-sized frame."""
    g = torch.Generator().manual_seed(seed)
    d0 = torch.nn.functional.normalize(torch.randn(m, 128, generator=g), dim=1)
    if n <= m:
        perm = torch.randperm(m, generator=g)[:n]
    else:
        perm = torch.cat([torch.randperm(m, generator=g), torch.randint(0, m, (n - m,), generator=g)])
    d1 = torch.nn.functional.normalize(d0[perm] + noise * torch.randn(n, 128, generator=g), dim=1)
    k0 = torch.rand(m, 2, generator=g) * torch.tensor([1241.0, 376.0])
    k1 = k0[perm] + torch.randn(n, 2, generator=g)
    return k0, d0, k1, d1, perm


def match_set(matches):
    a = matches.detach().cpu().numpy() if isinstance(matches, torch.Tensor) else np.asarray(matches)
    return set(map(tuple, a.reshape(-1, 2).tolist()))


def rel_err(got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-30))


def aliked_selection(score, n_limit, thr=0.2, radius=2):
    """The detector's discrete decisions restated on a score map [H,W] (numpy f32): 2-round NMS, border, threshold,
    top-n_limit by (score desc, raster index asc).  Returns (selected linear indices as a set, nms map, k-th score or None)."""
    import oracle
    nms = oracle.simple_nms(torch.from_numpy(score[None, None].copy()), radius)[0, 0].numpy().copy()
    nms[:radius] = 0; nms[-radius:] = 0; nms[:, :radius] = 0; nms[:, -radius:] = 0
    cand = np.flatnonzero(nms.ravel() > thr)
    kth = None
    if n_limit > 0 and len(cand) > n_limit:
        sc = score.ravel()[cand]
        order = np.argsort(-sc, kind="stable")
        kth = float(sc[order[n_limit - 1]])
        cand = cand[order[:n_limit]]
    return set(cand.tolist()), nms, kth


def aliked_decision_margins(score_gpu, score_ora, n_limit, thr=0.2):
    """discrete detector outputs (NMS equality, > 0.2, top-k cut) may flip on last-bit differences between
    two correct fp32 implementations.  For every pixel selected by exactly one of the two score maps, the DECISION MARGIN
    is the smallest of: |score - thr|, |score - k-th score| (when the list is truncated) and - for NMS ties - the score
    gap of any pixel pair (q, r), r in the 5x5 window of q, within 4 pixels of it whose ORDER differs between the two
    maps.  A flip with margin < 1e-6 is fp32 noise (reported, not hidden); anything larger is a real disagreement.
    Returns dict(n_diff, max_margin, margins=[(pixel, margin)], min_thr_margin, kth_gap)."""
    H, W = score_ora.shape
    sel_g, _, kth_g = aliked_selection(score_gpu, n_limit, thr)
    sel_o, nms_o, kth_o = aliked_selection(score_ora, n_limit, thr)
    out = []
    for p in sorted(sel_g ^ sel_o):
        y, x = divmod(p, W)
        m = min(abs(float(score_ora[y, x]) - thr), abs(float(score_gpu[y, x]) - thr))
        if kth_o is not None:
            m = min(m, abs(float(score_ora[y, x]) - kth_o), abs(float(score_gpu[y, x]) - kth_g))
        for qy in range(max(0, y - 4), min(H, y + 5)):
            for qx in range(max(0, x - 4), min(W, x + 5)):
                for ry in range(max(0, qy - 2), min(H, qy + 3)):
                    for rx in range(max(0, qx - 2), min(W, qx + 3)):
                        if (ry, rx) <= (qy, qx):
                            continue
                        do = float(score_ora[qy, qx]) - float(score_ora[ry, rx]); dg = float(score_gpu[qy, qx]) - float(score_gpu[ry, rx])
                        if (do > 0) != (dg > 0) or (do == 0) != (dg == 0):
                            m = min(m, max(abs(do), abs(dg)))
        out.append((p, m))
    cand = nms_o.ravel()[nms_o.ravel() > 0]
    srt = np.sort(score_ora.ravel()[np.flatnonzero(nms_o.ravel() > thr)])[::-1]
    gap = float(srt[n_limit - 1] - srt[n_limit]) if (n_limit > 0 and len(srt) > n_limit) else None
    return {"n_diff": len(out), "max_margin": max([m for _, m in out], default=0.0), "margins": out,
            "min_thr_margin": float(np.abs(cand - thr).min()) if len(cand) else None, "kth_gap": gap}


FM_CV_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fm_cv.npz")


def fm_cv_golden_cases():
    """(case, pts1, pts2, thresh, F, mask [n,1]) of the cv2-generated fixtures (tests/golden/make_golden_fm_cv.py)."""
    z = np.load(FM_CV_GOLD)
    c = 0
    while f"c{c}_cfg" in z:
        n = int(z[f"c{c}_cfg"][0])
        yield c, z[f"c{c}_pts1"], z[f"c{c}_pts2"], float(z[f"c{c}_cfg"][4]), z[f"c{c}_F"], np.unpackbits(z[f"c{c}_mask"])[:n].reshape(-1, 1)
        c += 1
