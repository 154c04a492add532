"""Stage-by-stage GPU-vs-oracle diagnostics (development aid; run under gpurun).
Writes a report to gpurun_out/diag.txt."""
import os, sys, time, traceback
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import weights, synth, frontend  # noqa: E402
import oracle  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n"); LOG.flush()


def cmp(name, got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    if got.shape != ref.shape:
        P(f"  {name}: SHAPE MISMATCH got {got.shape} ref {ref.shape}")
        return
    d = np.abs(got - ref)
    den = np.abs(ref).max() + 1e-30
    P(f"  {name}: shape {ref.shape} max_abs {d.max():.3e} rel_to_max {d.max()/den:.3e} mean_abs {d.mean():.3e} ref_absmax {den:.3e} nan {np.isnan(got).sum()}")


def aliked_diag(H=376, W=1241, max_kp=2048, model="aliked-n16"):
    P(f"=== ALIKED {model} {W}x{H} max_kp={max_kp}")
    sd = weights.synthetic_aliked_state(model)
    img = synth.frame(0, H, W)
    ora = oracle.ALIKED(model_name=model, max_num_keypoints=max_kp).eval()
    ora.load_state_dict(sd, strict=True)
    ora.record_taps = True
    t = time.time()
    fo = ora.extract(oracle.bgr_to_tensor(img))
    P(f"  oracle extract {time.time()-t:.2f}s  n={fo['keypoints'].shape[1]}")
    det = frontend.ALIKED(model_name=model, max_num_keypoints=max_kp, weights=sd)
    t = time.time()
    fg = det.extract_bgr(img)
    torch.cuda.synchronize()
    P(f"  gpu extract first call {time.time()-t:.3f}s  n={fg['keypoints'].shape[1]} launches={det.launches}")
    Hr, Wr, Hp, Wp = [int(v) for v in det.debug("geometry")]
    P(f"  geometry Hr={Hr} Wr={Wr} Hp={Hp} Wp={Wp}")
    T = ora.taps
    cmp("resized", det.debug("resized").reshape(3, Hr, Wr), T["resized"][0].numpy())
    cmp("padded", det.debug("padded").reshape(3, Hp, Wp), T["padded"][0].numpy())
    cmp("x1", det.debug("x1").reshape(16, Hp, Wp), T["x1"][0].numpy())
    cmp("x2", det.debug("x2").reshape(32, Hp // 2, Wp // 2), T["x2"][0].numpy())
    cmp("x3", det.debug("x3").reshape(Hp // 8, Wp // 8, 64).transpose(2, 0, 1), T["x3"][0].numpy())
    cmp("x4", det.debug("x4").reshape(Hp // 32, Wp // 32, 128).transpose(2, 0, 1), T["x4"][0].numpy())
    cmp("score_map", det.debug("score_map").reshape(Hr, Wr), T["score_map"][0, 0].numpy())
    cmp("feature_map", det.debug("feature_map").reshape(Hr, Wr, 128).transpose(2, 0, 1), T["feature_map"][0].numpy())
    nms_o = oracle.simple_nms(T["score_map"], 2)[0, 0].numpy().copy()
    nms_o[:2] = 0; nms_o[-2:] = 0; nms_o[:, :2] = 0; nms_o[:, -2:] = 0
    nms_g = det.debug("nms").reshape(Hr, Wr)
    P(f"  nms support: oracle {int((nms_o>0).sum())} gpu {int((nms_g>0).sum())} differing pixels {int(((nms_o>0)!=(nms_g>0)).sum())}")
    # NMS of the GPU's own score map by the oracle (isolates the NMS kernel from conv rounding)
    sg = torch.from_numpy(det.debug("score_map").reshape(1, 1, Hr, Wr).copy())
    nms_s = oracle.simple_nms(sg, 2)[0, 0].numpy().copy()
    nms_s[:2] = 0; nms_s[-2:] = 0; nms_s[:, :2] = 0; nms_s[:, -2:] = 0
    P(f"  nms kernel vs oracle-nms(gpu score): differing pixels {int((nms_s != nms_g).sum())}")
    n = fg["keypoints"].shape[1]
    ko, kg = fo["keypoints"][0].numpy(), fg["keypoints"][0].cpu().numpy()
    io = T["dkd_indices"].numpy()
    kn = det.debug("kp_norm").reshape(-1, 2)[:n]
    # integer pixel index of GPU keypoints: recover from kp_norm (rounded)
    xg = np.rint((kn[:, 0] + 1) / 2 * (Wr - 1)); yg = np.rint((kn[:, 1] + 1) / 2 * (Hr - 1))
    so, sg_ = set(io.tolist()), None
    P(f"  keypoints: oracle n={len(io)} gpu n={n}")
    if n == len(io):
        same_order = 0
        # compare order: positions close
        dd = np.abs(ko - kg).max(axis=1)
        P(f"  same-order fraction (|dkp|<1e-2 px): {(dd < 1e-2).mean():.4f}; max |dkp| where same {dd[dd<1e-2].max() if (dd<1e-2).any() else -1:.3e}")
    # set comparison by nearest pixel
    from collections import Counter
    go = Counter((int(round(x)), int(round(y))) for x, y in ko)
    gg = Counter((int(round(x)), int(round(y))) for x, y in kg)
    P(f"  keypoint set (rounded orig px): common {sum((go & gg).values())} only_oracle {sum((go - gg).values())} only_gpu {sum((gg - go).values())}")
    if n == len(io) and (np.abs(ko - kg).max(axis=1) < 1e-2).mean() > 0.99:
        cmp("descriptors", fg["descriptors"][0].cpu().numpy(), fo["descriptors"][0].numpy())
        cmp("keypoint_scores", fg["keypoint_scores"][0].cpu().numpy(), fo["keypoint_scores"][0].numpy())
        cmp("sddh_offset", det.debug("sddh_offset").reshape(-1, det._handle and (32 if model.endswith('n32') else 16), 2)[:n], T["sddh_offset"].numpy())
        cmp("desc_raw", det.debug("desc_raw").reshape(-1, 128)[:n], T["sddh_desc_raw"].numpy())
    # host API
    kp_h, de_h, sc_h = det.extract_host(img)
    P(f"  extract_host n={len(kp_h)} max|kp diff vs device path| {np.abs(kp_h - kg).max() if len(kp_h)==n else 'n/a'}")
    # timing
    t0 = torch.from_numpy(img).cuda()
    for _ in range(3):
        det.extract_device(t0, 0, H, W, 3 * W)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        det.extract_device(t0, 0, H, W, 3 * W)
    ev1.record(); torch.cuda.synchronize()
    P(f"  extract device time: {ev0.elapsed_time(ev1)/10:.3f} ms/frame")
    return fo, fg, sd


def lg_diag(m=2048, n=2048, seed=0, final_scale=16.0, label="default", **lgkw):
    P(f"=== LightGlue m={m} n={n} [{label}] {lgkw}")
    g = torch.Generator().manual_seed(seed)
    d0 = torch.nn.functional.normalize(torch.randn(m, 128, generator=g), dim=1)
    perm = torch.randperm(m, generator=g)[:n] if n <= m else torch.cat([torch.randperm(m, generator=g), torch.randint(0, m, (n - m,), generator=g)])
    d1 = torch.nn.functional.normalize(d0[perm] + 0.05 * torch.randn(n, 128, generator=g), dim=1)
    k0 = torch.rand(m, 2, generator=g) * torch.tensor([1241.0, 376.0])
    k1 = k0[perm] + torch.randn(n, 2, generator=g)
    sd = weights.synthetic_lightglue_state(seed=seed, final_scale=final_scale, **{k: v for k, v in lgkw.items() if k in ("token_bias", "token_gain", "match_bias", "match_gain")})
    okw = {k: v for k, v in lgkw.items() if k in ("depth_confidence", "width_confidence", "filter_threshold")}
    ora = oracle.LightGlue(**okw).eval()
    missing = ora.load_state_dict(sd, strict=False)
    ora.record_taps = True
    t = time.time()
    ro = ora({"image0": {"keypoints": k0[None], "descriptors": d0[None]}, "image1": {"keypoints": k1[None], "descriptors": d1[None]}})
    P(f"  oracle {time.time()-t:.2f}s matches={len(ro['matches'][0])} stop={ro['stop']} prune0 hist={np.bincount(ro['prune0'][0].long().numpy().astype(int)).tolist()}")
    mat = frontend.LightGlue(weights=sd, **okw)
    mat.set_debug(True)
    t = time.time()
    rg = mat({"image0": {"keypoints": k0[None].cuda(), "descriptors": d0[None].cuda()}, "image1": {"keypoints": k1[None].cuda(), "descriptors": d1[None].cuda()}})
    P(f"  gpu first call {time.time()-t:.3f}s matches={len(rg['matches'][0])} stop={rg['stop']} launches={mat.launches}")
    L = ora.taps["layers"]
    for i, (a, b) in enumerate(L):
        try:
            g0 = mat.debug(f"layer{i}_0").reshape(-1, 256); g1 = mat.debug(f"layer{i}_1").reshape(-1, 256)
            cmp(f"layer{i} desc0", g0, a[0].numpy()); cmp(f"layer{i} desc1", g1, b[0].numpy())
        except Exception as e:
            P(f"  layer{i}: {e}")
    try:
        sim_g = mat.debug("sim")
        cmp("sim", sim_g.reshape(ora.taps["sim"].shape[1:]), ora.taps["sim"][0].numpy())
    except Exception as e:
        P(f"  sim: {e}")
    mo = ro["matches"][0].numpy(); mg = rg["matches"][0].cpu().numpy()
    so = set(map(tuple, mo.tolist())); sg = set(map(tuple, mg.tolist()))
    P(f"  match sets: oracle {len(so)} gpu {len(sg)} common {len(so & sg)} identical={so == sg} same_order={mo.shape == mg.shape and (mo == mg).all()}")
    if mo.shape == mg.shape and (mo == mg).all():
        cmp("scores", rg["scores"][0].cpu().numpy(), ro["scores"][0].numpy())
    for k in ("matches0", "matches1", "prune0", "prune1"):
        a, b = ro[k][0].long().numpy(), rg[k][0].cpu().long().numpy()
        P(f"  {k}: equal={np.array_equal(a, b)} ndiff={(a != b).sum() if a.shape == b.shape else 'shape'}")
    for k in ("matching_scores0", "matching_scores1"):
        cmp(k, rg[k][0].cpu().numpy(), ro[k][0].numpy())
    # timing of the device path
    dk0, dd0, dk1, dd1 = k0.cuda(), d0.cuda(), k1.cuda(), d1.cuda()
    mat.set_debug(False)
    for _ in range(2):
        mat.match_device(dk0, dd0, dk1, dd1)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5):
        mat.match_device(dk0, dd0, dk1, dd1)
    ev1.record(); torch.cuda.synchronize()
    P(f"  match device time: {ev0.elapsed_time(ev1)/5:.3f} ms/pair")


def main():
    P(torch.cuda.get_device_name(0), torch.version.cuda)
    steps = [
        lambda: lg_diag(256, 200, label="small ragged"),
        lambda: lg_diag(2048, 2048, label="headline"),
        lambda: lg_diag(1024, 900, label="adaptive", token_bias=1.4, token_gain=6.0, match_bias=-5.0, match_gain=4.0, filter_threshold=1e-6),
        lambda: lg_diag(512, 512, label="adaptive-nostop", token_bias=1.0, token_gain=6.0, match_bias=-5.0, match_gain=4.0, filter_threshold=1e-6),
        lambda: lg_diag(512, 512, label="no-adaptive", depth_confidence=-1, width_confidence=-1),
        lambda: aliked_diag(376, 1241, 2048),
        lambda: aliked_diag(480, 640, 1024),
        lambda: aliked_diag(200, 200, 4000),
    ]
    for s in steps:
        try:
            s()
        except Exception:
            P("EXCEPTION:\n" + traceback.format_exc())
    P("done")


if __name__ == "__main__":
    main()
