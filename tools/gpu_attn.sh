#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/trace_attn.py 2>&1 | tee gpurun_out/trace_attn.txt | tail -8
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_lightglue.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_perf.log
timeout 200 python tools/time_stages.py fp32 2>&1 | tee gpurun_out/time_stages.txt
