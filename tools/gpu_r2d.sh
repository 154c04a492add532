#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_tensorcore.py -x -q -k "gemm" > gpurun_out/r2d_pytest_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest_gemm.log; tail -5 gpurun_out/r2d_pytest_gemm.log
grep -q "rc=0" gpurun_out/r2d_pytest_gemm.log || exit 1
timeout 200 python tools/bench_gemm2.py 2>&1 | tee gpurun_out/r2d_bench_gemm2.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_gpu.log
tail -12 gpurun_out/r2d_pytest_gpu.log
timeout 300 python tools/time_batch.py fp32,bf16 1,8 2>/dev/null | tee gpurun_out/r2d_time_batch.txt
B2S_GEMM_PERSIST=0 timeout 300 python tools/time_batch.py fp32,bf16 8 2>/dev/null | tee gpurun_out/r2d_time_batch_nopersist.txt
