#!/bin/bash
# One GPU-box pass: full parity suite, geometry micro-bench, smoke, headline bench (fp32 incl. CPU baseline).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tests/perf_geometry.py > gpurun_out/bench_geometry.jsonl 2> gpurun_out/bench_geometry.err; cat gpurun_out/bench_geometry.jsonl; tail -3 gpurun_out/bench_geometry.err
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -c 400 gpurun_out/bench_fp32.err; cut -c 1-400 gpurun_out/bench_fp32.json
