#!/bin/bash
# round-2 check of the batched matcher: parity suite, then batch-size sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
tail -25 gpurun_out/r2b_pytest_gpu.log
timeout 600 python tools/time_batch.py fp32,bf16 1,2,4,8,16 > gpurun_out/r2b_time_batch.txt 2>&1; cat gpurun_out/r2b_time_batch.txt | tail -30
