"""GPU: unit tests of the tcgen05/TMEM kernels (bf16 operands, fp32 accumulation) and the bf16
LightGlue path.  Tolerances: GEMM/attention vs a torch fp32 reference on the SAME bf16-rounded
operands: 2e-3 relative to the output scale; bf16 matcher vs fp32 oracle: match-set agreement
>= 99% (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle
from b200slam import weights
from helpers import noisy_copy_pair, match_set

pytestmark = pytest.mark.gpu


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 256), (200, 256, 256), (4096, 768, 256), (1000, 512, 512), (333, 256, 512)])
def test_gemm_tc(M, N, K):
    from b200slam import _lib
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    out = np.empty((M, N), np.float32)
    _lib.check(_lib.lib.b2s_test_gemm_tc(A.numpy().ctypes.data, W.numpy().ctypes.data, b.numpy().ctypes.data, M, N, K, out.ctypes.data), "gemm_tc")
    ref = (_bf16(A) @ _bf16(W).T + b).numpy()
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert err < 2e-3, f"rel err {err}"


@pytest.mark.parametrize("nq,nk", [(128, 128), (128, 256), (300, 200), (2048, 2048), (1, 77), (129, 1)])
def test_attn_tc(nq, nk):
    from b200slam import _lib
    g = torch.Generator().manual_seed(nq * 7 + nk)
    q = torch.randn(nq, 256, generator=g); k = torch.randn(nk, 256, generator=g); v = torch.randn(nk, 256, generator=g)
    out = np.empty((nq, 256), np.float32)
    _lib.check(_lib.lib.b2s_test_attn_tc(q.numpy().ctypes.data, k.numpy().ctypes.data, v.numpy().ctypes.data, nq, nk, out.ctypes.data), "attn_tc")
    sp = lambda t: _bf16(t).view(-1, 4, 64).transpose(0, 1)   # noqa: E731
    ref = torch.nn.functional.scaled_dot_product_attention(sp(q)[None], sp(k)[None], sp(v)[None])[0].transpose(0, 1).reshape(nq, 256).numpy()
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert err < 2e-2, f"rel err {err}"      # P and the output are rounded to bf16 (2^-8)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (200, 256, 256), (4096, 768, 256), (1000, 512, 512), (333, 256, 512)])
def test_gemm_tc3_is_fp32_faithful(M, N, K):
    """fp32 operands carried as three bf16 planes, six cross products: error vs a float64 reference must be at the
    level of an fp32 GEMM: tolerance 1e-6 of the output scale and no worse than 2x torch's own fp32 matmul + 1e-7.
    (The operand split is exact to 2^-26; the tensor core truncates its fp32 accumulator after every k-step, which
    the kernel contains by spreading the accumulation over several TMEM accumulators summed in registers.)"""
    from b200slam import _lib
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    out = np.empty((M, N), np.float32)
    _lib.check(_lib.lib.b2s_test_gemm_tc3(A.numpy().ctypes.data, W.numpy().ctypes.data, b.numpy().ctypes.data, M, N, K, out.ctypes.data), "gemm_tc3")
    ref = (A.double() @ W.double().T + b.double()).numpy()
    err = np.abs(out - ref).max() / np.abs(ref).max()
    err32 = np.abs((A @ W.T + b).numpy() - ref).max() / np.abs(ref).max()
    assert err < 1e-6 and err < 2 * err32 + 1e-7, f"rel err {err} (torch fp32: {err32})"


@pytest.mark.parametrize("nq,nk", [(128, 128), (128, 64), (300, 200), (2048, 2048), (1, 77), (129, 1), (640, 1000)])
def test_attn_tc3_is_fp32_faithful(nq, nk):
    from b200slam import _lib
    g = torch.Generator().manual_seed(nq * 7 + nk)
    q = 2 * torch.randn(nq, 256, generator=g); k = torch.randn(nk, 256, generator=g); v = torch.randn(nk, 256, generator=g)
    out = np.empty((nq, 256), np.float32)
    _lib.check(_lib.lib.b2s_test_attn_tc3(q.numpy().ctypes.data, k.numpy().ctypes.data, v.numpy().ctypes.data, nq, nk, out.ctypes.data), "attn_tc3")
    sp = lambda t: t.double().view(-1, 4, 64).transpose(0, 1)   # noqa: E731
    ref = torch.nn.functional.scaled_dot_product_attention(sp(q)[None], sp(k)[None], sp(v)[None])[0].transpose(0, 1).reshape(nq, 256).numpy()
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert err < 5e-6, f"rel err {err}"      # fp32 softmax with ex2.approx: a few fp32 ulps of the output scale


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (200, 256, 256), (4096, 768, 256), (1000, 512, 512), (333, 256, 512)])
@pytest.mark.parametrize("scale", [1.0, 1e-3, 300.0])
def test_gemm_h2_is_fp32_faithful(M, N, K, scale):
    """fp32 operands carried as two fp16 planes (x = h0 + 2^-11 h1), three cross products, main / correction terms in
    separate accumulators: 1e-6 of the output scale and no worse than 3x torch's own fp32 matmul + 1e-7 (measured on
    B200: 4.6e-7 at K = 256, 7.4 - 8.6e-7 at K = 512; bf16x3: 2.9e-7 / 5.5e-7; torch fp32: 5.9e-7 / 3.4 - 4.0e-7 - the
    tensor core's per-k-step accumulator truncation, not the operand split, dominates both), at activation magnitudes
    from 1e-3 (residual planes would be fp16-subnormal without the 2^11 residual scaling) to 300."""
    from b200slam import _lib
    g = torch.Generator().manual_seed(M + N + K)
    A = scale * torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = scale * torch.randn(N, generator=g)
    out = np.empty((M, N), np.float32)
    _lib.check(_lib.lib.b2s_test_gemm_h2(A.numpy().ctypes.data, W.numpy().ctypes.data, b.numpy().ctypes.data, M, N, K, out.ctypes.data), "gemm_h2")
    ref = (A.double() @ W.double().T + b.double()).numpy()
    err = np.abs(out - ref).max() / np.abs(ref).max()
    err32 = np.abs((A @ W.T + b).numpy() - ref).max() / np.abs(ref).max()
    assert err < 1e-6 and err < 3 * err32 + 1e-7, f"rel err {err} (torch fp32: {err32})"


@pytest.mark.parametrize("nq,nk", [(128, 128), (128, 64), (300, 200), (2048, 2048), (1, 77), (129, 1), (640, 1000)])
def test_attn_h2_is_fp32_faithful(nq, nk):
    from b200slam import _lib
    g = torch.Generator().manual_seed(nq * 7 + nk)
    q = 2 * torch.randn(nq, 256, generator=g); k = torch.randn(nk, 256, generator=g); v = torch.randn(nk, 256, generator=g)
    out = np.empty((nq, 256), np.float32)
    _lib.check(_lib.lib.b2s_test_attn_h2(q.numpy().ctypes.data, k.numpy().ctypes.data, v.numpy().ctypes.data, nq, nk, out.ctypes.data), "attn_h2")
    sp = lambda t: t.double().view(-1, 4, 64).transpose(0, 1)   # noqa: E731
    ref = torch.nn.functional.scaled_dot_product_attention(sp(q)[None], sp(k)[None], sp(v)[None])[0].transpose(0, 1).reshape(nq, 256).numpy()
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert err < 5e-6, f"rel err {err}"      # same bar as the bf16x3 kernel


def test_fp16_range_overflow_falls_back_to_bf16x3():
    """precision='fp32' (fp16x2 planes): descriptors beyond the fp16 range raise the device flag, the pair reports
    LG_RANGE and is re-run on the bf16x3 engine - results equal a precision='fp32x3' matcher bit for bit."""
    from b200slam import frontend
    sd = weights.synthetic_lightglue_state(seed=0)
    k0, d0, k1, d1, _ = noisy_copy_pair(300, 280, seed=3)
    d0 = d0.clone(); d0[7, 5] = 1.0e5           # > 65504
    fast = frontend.LightGlue(weights=sd, precision="fp32", max_kp=512)
    safe = frontend.LightGlue(weights=sd, precision="fp32x3", max_kp=512)
    assert fast.planes == 2 and safe.planes == 3
    rs = safe.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy(), full=True)
    rf = fast.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy(), full=True)
    assert int(_lib_fallbacks(fast)) == 1
    for k in rs:
        assert np.array_equal(np.asarray(rf[k]), np.asarray(rs[k])), k
    # the batched device path reports LG_RANGE and resolve_range repairs it; the flag does not stick to later batches
    out = fast.match_batch_device([k0.cuda(), k1.cuda()], [d0.cuda(), d1.cuda()], [(0, 1), (1, 1)])
    n = out["n"].cpu().numpy()
    assert n[0] == len(rs["matches"]) and n[1] > 0
    k0b, d0b, k1b, d1b, _ = noisy_copy_pair(300, 280, seed=3)
    r2 = fast.match_host(k0b.numpy(), d0b.numpy(), k1b.numpy(), d1b.numpy())
    assert int(_lib_fallbacks(fast)) == 1 + 1 and len(r2["matches"]) > 20     # one more from the batch, none from the clean pair


def _lib_fallbacks(mat):
    from b200slam import _lib
    return _lib.lib.b2s_lg_range_fallbacks(mat._handle) + getattr(mat, "range_fallbacks", 0)


@pytest.mark.parametrize("m,n", [(2048, 2048), (700, 512)])
def test_bf16_matcher_agreement_with_fp32_oracle(m, n):
    from b200slam import frontend
    sd = weights.synthetic_lightglue_state(seed=0)
    ora = oracle.LightGlue().eval(); ora.load_state_dict(sd, strict=False); ora.record_taps = True
    mat = frontend.LightGlue(weights=sd, precision="bf16", max_kp=max(m, n))
    mat.set_debug(True)
    k0, d0, k1, d1, _ = noisy_copy_pair(m, n, seed=5)
    ro = ora({"image0": {"keypoints": k0[None], "descriptors": d0[None]}, "image1": {"keypoints": k1[None], "descriptors": d1[None]}})
    rg = mat.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy())
    for i, (a, b) in enumerate(ora.taps["layers"]):
        e = np.abs(mat.debug(f"layer{i}_0").reshape(-1, 256) - a[0].numpy()).max() / np.abs(a[0].numpy()).max()
        assert e < 5e-2, f"layer {i}: rel err {e}"
    so, sg = match_set(ro["matches"][0]), match_set(rg["matches"])
    agree = len(so & sg) / max(len(so | sg), 1)
    assert len(so) > 300 and agree >= 0.99, f"match-set agreement {agree:.4f} ({len(so)} vs {len(sg)})"
