#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_gpu.log
tail -6 gpurun_out/r2h_pytest_gpu.log
timeout 300 python tools/profile_e2e.py fp32 2>/dev/null | tee gpurun_out/r2h_profile_e2e.txt
timeout 300 python tools/report_margins.py 4 2>/dev/null > gpurun_out/r2h_decision_margins.json; head -c 1500 gpurun_out/r2h_decision_margins.json
