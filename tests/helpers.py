"""Shared test inputs (seeded, regenerated - never read from /root/reference)."""
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
