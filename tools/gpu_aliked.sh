#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_aliked.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_perf.log
timeout 200 python tools/time_stages.py none 2>&1 | tee gpurun_out/time_stages.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_aliked.csv python tools/prof_aliked.py 2 > gpurun_out/ncu_aliked_l.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_aliked.csv | head -16
