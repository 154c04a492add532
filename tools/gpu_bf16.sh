#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_lightglue.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_perf.log
timeout 200 python tools/time_stages.py fp32,bf16 2>&1 | tee gpurun_out/time_stages.txt
timeout 100 python tools/bench_kernels.py attn 2>&1 | tail -6
