#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_gpu.log
tail -12 gpurun_out/r2c_pytest_gpu.log
for b in 1 8; do for p in fp32 bf16; do timeout 120 python tools/prof_batch.py $p $b 4 1 2>/dev/null | tee -a gpurun_out/r2c_prof_classes.txt; done; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 115 -c 114 --csv --log-file gpurun_out/r2c_launches_fp32_b8.csv python tools/prof_batch.py fp32 8 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 115 -c 114 --csv --log-file gpurun_out/r2c_launches_bf16_b8.csv python tools/prof_batch.py bf16 8 1 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2c_launches_fp32_b8.csv 2>/dev/null | head -30
python tools/summarize_launches.py gpurun_out/r2c_launches_bf16_b8.csv 2>/dev/null | head -30
