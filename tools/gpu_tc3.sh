#!/bin/bash
# first GPU pass of the fp32-on-bf16x3 tensor-core path: unit tests, matcher parity, timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_lightglue.py -m gpu -q 2>&1 | tail -15
timeout 300 python tools/bench_kernels.py attn 2>&1 | tee gpurun_out/bench_attn.txt
timeout 600 python tools/time_stages.py 2>&1 | tee gpurun_out/time_stages.txt
