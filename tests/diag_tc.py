"""tcgen05 kernel bring-up diagnostics (run under gpurun with a timeout)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from b200slam import _lib, weights, frontend
import oracle
from helpers import noisy_copy_pair, match_set

def bf(x): return x.to(torch.bfloat16).to(torch.float32)

def gemm(M, N, K):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    out = np.full((M, N), np.nan, np.float32)
    rc = _lib.lib.b2s_test_gemm_tc(A.numpy().ctypes.data, W.numpy().ctypes.data, b.numpy().ctypes.data, M, N, K, out.ctypes.data)
    ref = (bf(A) @ bf(W).T + b).numpy()
    err = np.abs(out - ref)
    print(f"gemm {M}x{N}x{K}: rc={rc} {_lib.lib.b2s_last_error().decode() if rc else ''} max_abs={np.nanmax(err):.4e} ref_max={np.abs(ref).max():.3f} nan={np.isnan(out).sum()}", flush=True)
    if np.nanmax(err) > 0.05:
        bad = np.argwhere(err > 0.05)
        print("   bad rows (first 10):", sorted(set(bad[:, 0].tolist()))[:10], "bad cols:", sorted(set(bad[:, 1].tolist()))[:10], "n_bad", len(bad))
        print("   out[0,:8]", out[0, :8], "\n   ref[0,:8]", ref[0, :8])

def attn(nq, nk):
    g = torch.Generator().manual_seed(2)
    q = torch.randn(nq, 256, generator=g); k = torch.randn(nk, 256, generator=g); v = torch.randn(nk, 256, generator=g)
    out = np.full((nq, 256), np.nan, np.float32)
    rc = _lib.lib.b2s_test_attn_tc(q.numpy().ctypes.data, k.numpy().ctypes.data, v.numpy().ctypes.data, nq, nk, out.ctypes.data)
    sp = lambda t: bf(t).view(-1, 4, 64).transpose(0, 1)
    ref = torch.nn.functional.scaled_dot_product_attention(sp(q)[None], sp(k)[None], sp(v)[None])[0].transpose(0, 1).reshape(nq, 256).numpy()
    err = np.abs(out - ref)
    print(f"attn {nq}x{nk}: rc={rc} {_lib.lib.b2s_last_error().decode() if rc else ''} max_abs={np.nanmax(err):.4e} ref_max={np.abs(ref).max():.3f} nan={np.isnan(out).sum()}", flush=True)
    if np.nanmax(err) > 0.05:
        print("   out[0,:8]", out[0, :8], "\n   ref[0,:8]", ref[0, :8])

def matcher(m, n):
    sd = weights.synthetic_lightglue_state(seed=0)
    ora = oracle.LightGlue().eval(); ora.load_state_dict(sd, strict=False); ora.record_taps = True
    mat = frontend.LightGlue(weights=sd, precision="bf16", max_kp=max(m, n)); mat.set_debug(True)
    k0, d0, k1, d1, _ = noisy_copy_pair(m, n, seed=5)
    ro = ora({"image0": {"keypoints": k0[None], "descriptors": d0[None]}, "image1": {"keypoints": k1[None], "descriptors": d1[None]}})
    rg = mat.match_host(k0.numpy(), d0.numpy(), k1.numpy(), d1.numpy())
    for i, (a, b) in enumerate(ora.taps["layers"]):
        e = np.abs(mat.debug(f"layer{i}_0").reshape(-1, 256) - a[0].numpy()).max() / np.abs(a[0].numpy()).max()
        print(f"  layer{i} rel err {e:.3e}")
    so, sg = match_set(ro["matches"][0]), match_set(rg["matches"])
    print(f"matcher bf16 {m}x{n}: oracle {len(so)} gpu {len(sg)} common {len(so & sg)} agreement {len(so & sg)/max(len(so | sg),1):.4f}", flush=True)
    mat.set_debug(False)
    dk0, dd0, dk1, dd1 = k0.cuda(), d0.cuda(), k1.cuda(), d1.cuda()
    for _ in range(3): mat.match_device(dk0, dd0, dk1, dd1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): mat.match_device(dk0, dd0, dk1, dd1)
    e1.record(); torch.cuda.synchronize()
    print(f"  bf16 match device time {e0.elapsed_time(e1)/10:.3f} ms/pair", flush=True)

if __name__ == "__main__":
    what = sys.argv[1]
    if what == "gemm":
        for s in [(128, 64, 64), (128, 128, 64), (128, 128, 256), (256, 256, 256), (4096, 768, 256), (1000, 512, 512), (333, 256, 512)]: gemm(*s)
    elif what == "attn":
        for s in [(128, 128), (128, 256), (300, 200), (2048, 2048), (1, 77)]: attn(*s)
    elif what == "matcher":
        matcher(2048, 2048); matcher(700, 512)
