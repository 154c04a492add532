"""GPU vs cv2 timing of the rows behind the matcher: F-matrix RANSAC per call (device-resident and host buffers) and
u8 remap per frame.  Prints one JSON line per row; the numbers under profiles/ come from this script."""
import json
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from b200slam.geometry import FrameUndistorter, FundamentalRansac   # noqa: E402
from b200slam import synth                                          # noqa: E402
from oracle import geometry as G                                    # noqa: E402  (scene generator only)


def cuda_ms(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def wall_ms(fn, iters=20, warm=2):
    for _ in range(warm):
        fn()
    t = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t) * 1e3 / iters


def main():
    r = FundamentalRansac(n_hyp=2048)
    for n, frac in [(700, 0.3), (1300, 0.3), (2048, 0.5)]:
        p1, p2, gt = G.two_view_scene(n, frac, 0.2, 1)
        d1, d2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
        dev = cuda_ms(lambda: r.run_device(d1, d2, None, n, 1.0))
        host = wall_ms(lambda: r.run_host(p1, p2, 1.0))
        cv = wall_ms(lambda: cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99), iters=10)
        _, m = r.run_host(p1, p2, 1.0)
        _, mc = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
        # fp64 work of the scoring kernel: 2048 samples x (<=3 models) x n points x ~45 flop
        print(json.dumps({"row": "f1 fundamental-matrix RANSAC", "n_matches": n, "outlier_frac": frac, "n_hyp": 2048,
                          "gpu_device_ms": round(dev, 4), "gpu_host_buffers_ms": round(host, 4), "cv2_cpu_ms": round(cv, 3),
                          "speedup_host_api": round(cv / host, 1), "inliers_gpu": int(m.sum()), "inliers_cv2": int(mc.sum()),
                          "true_inliers": int(gt.sum())}))
    K = np.array([[718.856, 0, 607.19], [0, 718.856, 185.2], [0, 0, 1.0]]); D = np.array([-0.28, 0.07, 0.0002, 0.0001, 0.0])
    nK, _ = cv2.getOptimalNewCameraMatrix(K, D, (1241, 376), alpha=0, newImgSize=(1241, 376))
    mx, my = cv2.initUndistortRectifyMap(K, D, None, nK, (1241, 376), cv2.CV_32FC1)
    und = FrameUndistorter(mx, my)
    img = synth.frame(0, 376, 1241)
    src = torch.from_numpy(img).cuda(); dst = torch.empty_like(src)
    from b200slam._lib import lib
    st = torch.cuda.current_stream().cuda_stream
    dev = cuda_ms(lambda: lib.b2s_remap_bgr(und._handle, src.data_ptr(), 3 * 1241, st, dst.data_ptr(), 3 * 1241), iters=200)
    host = wall_ms(lambda: und.remap(img))
    cv = wall_ms(lambda: cv2.remap(img, mx, my, cv2.INTER_LINEAR))
    bytes_alg = img.size * 2 + mx.size * 8      # u8 in + u8 out + two f32 maps
    print(json.dumps({"row": "f4 frame ingest (cv2.remap u8 BGR 1241x376)", "gpu_device_ms": round(dev, 4),
                      "gpu_host_buffers_ms": round(host, 4), "cv2_cpu_ms": round(cv, 3), "algorithmic_bytes": bytes_alg,
                      "achieved_GBps": round(bytes_alg / dev / 1e6, 1), "bit_exact": bool(np.array_equal(und.remap(img), cv2.remap(img, mx, my, cv2.INTER_LINEAR)))}))




def bench_reproj():
    from b200slam import pnp_utils as P
    from oracle import pnp as O
    for n_points in (1500, 6000):
        wm, K, Tcw, kps, des = O.tracking_scene(n_points, 2048, 0)
        m = P.ReprojectionMatcher()
        r = m.match(wm, K, Tcw, kps, des, 1241, 376, 12.0, 0.8)          # first call uploads the whole mirror
        warm = wall_ms(lambda: m.match(wm, K, Tcw, kps, des, 1241, 376, 12.0, 0.8), iters=10)
        t = time.perf_counter(); o = O.reproject_and_match_2d3d(wm, K, Tcw, kps, des, 1241, 376, 12.0, 0.8); cpu = (time.perf_counter() - t) * 1e3
        # device-only: the two kernels on resident buffers
        mir = m.mirror_of(wm); ids, pos, rows = mir.sync(wm)
        Xw = torch.from_numpy(pos).cuda(); rd = torch.from_numpy(rows).cuda(); kd = torch.from_numpy(kps).cuda(); dd = torch.from_numpy(des).cuda()
        out = torch.empty((len(ids),), dtype=torch.int32, device="cuda"); Kh = np.ascontiguousarray(K, np.float64).reshape(9); Th = np.ascontiguousarray(Tcw, np.float64).reshape(16)
        from b200slam._lib import lib
        st = torch.cuda.current_stream().cuda_stream
        dev = cuda_ms(lambda: lib.b2s_reproj_match(m._handle, Xw.data_ptr(), mir.desc.data_ptr(), mir.nobs.data_ptr(), rd.data_ptr(), len(ids), 6,
                                                   Kh.ctypes.data, Th.ctypes.data, kd.data_ptr(), dd.data_ptr(), len(kps), 1241, 376, 12.0, 0.8, st,
                                                   out.data_ptr(), None, None), iters=100)
        print(json.dumps({"row": "f3 reproject_and_match_2d3d", "map_points": n_points, "keypoints": 2048, "matches": len(r.mp_ids),
                          "identical_to_cpu_restatement": r.mp_ids == o.mp_ids and r.kp_indices == o.kp_indices,
                          "gpu_device_ms": round(dev, 4), "gpu_drop_in_ms": round(warm, 3), "cpu_restatement_ms": round(cpu, 1)}))


if __name__ == "__main__":
    main()
    bench_reproj()
