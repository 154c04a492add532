"""BASELINE config 3 on N GPUs: 16 keyframes -> 16 extractions + one feature all_gather + 120 pair matches.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_window.py [--precision fp32]
Prints one JSON line (rank 0): window time = max over ranks of the device time (CUDA events)."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200slam import frontend, synth, weights, window

ap = argparse.ArgumentParser(); ap.add_argument("--precision", default="fp32"); ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
H, W, NKP, NKF = 376, 1241, 2048, 16
sa, _ = weights.load_aliked_state(allow_synthetic=True); sl, _ = weights.load_lightglue_state(allow_synthetic=True)
det = frontend.ALIKED(max_num_keypoints=NKP, weights=sa, device=dev)
mat = frontend.LightGlue(weights=sl, device=dev, precision=args.precision, max_kp=NKP)
frames = [torch.from_numpy(synth.frame(8 * t, H, W)).to(dev) for t in range(NKF)]     # keyframes t = 0, 8, ..., 120
times, nm = [], 0
for rep in range(args.reps + 1):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = window.match_keyframe_window(frames, det, mat, H, W, rank, world)
    e1.record(); torch.cuda.synchronize()
    if rep:
        times.append(e0.elapsed_time(e1))
    nm = sum(int(v[2]) for v in res.values())
t = torch.tensor([sum(times) / len(times), float(nm), float(len(res))], dtype=torch.float64, device=dev)
tmax = t.clone()
if world > 1:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(t, op=dist.ReduceOp.SUM)
if rank == 0:
    pairs = int(t[2]) if world > 1 else len(res)
    print(json.dumps({"workload": "keyframe window all-pairs (BASELINE config 3): 16 keyframes, 120 pairs, 2048 kp", "n_gpus": world,
                      "precision": args.precision, "window_ms": float(tmax[0]), "pairs": pairs, "pairs_per_s": pairs / (float(tmax[0]) * 1e-3),
                      "total_matches": int(t[1]) if world > 1 else nm,
                      "collective": "one all_gather of (count, kpts, desc) records, 17 MB total" if world > 1 else "none"}))
if world > 1:
    dist.destroy_process_group()
