#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_geometry.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_geometry.log
timeout 300 python tests/perf_geometry.py > gpurun_out/bench_geometry.jsonl 2> gpurun_out/bench_geometry.err; cat gpurun_out/bench_geometry.jsonl; tail -5 gpurun_out/bench_geometry.err
