#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest_gpu.log
tail -15 gpurun_out/r2m_pytest_gpu.log
python tools/sanitize_target.py gemm 2>/dev/null | grep tc3
timeout 300 python tools/time_batch.py fp32,bf16 1,8 2>/dev/null | grep -v full-depth | tee gpurun_out/r2m_time_batch.txt
B2S_GEMM_WIDE=0 timeout 300 python tools/time_batch.py fp32 8 2>/dev/null | grep adaptive
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py batch > gpurun_out/r2_sanitizer_memcheck_batch.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r2_sanitizer_memcheck_batch.log
