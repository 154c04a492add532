#!/bin/bash
# multi-GPU bench exactly as the driver launches it:  tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "rc=$?"; tail -c 800 gpurun_out/r2_bench_${N}gpu.err | grep -v Warning | tail -5
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print({k:d[k] for k in ["value","n_gpus","ms_per_step"]}); print("e2e", d["e2e"]["value"]); print(d.get("config3_window")); print(d.get("other_precision",{}).get("value")); print(d["clocks"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c 1-300
