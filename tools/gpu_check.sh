#!/bin/bash
# One GPU-box pass: parity tests, both bench precisions, ncu launch list + full capture of the top kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -c 600 gpurun_out/bench_bf16.err
timeout 600 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -c 600 gpurun_out/bench_fp32.err
cat gpurun_out/bench_bf16.json gpurun_out/bench_fp32.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 600 --csv --log-file gpurun_out/launches_bf16.csv python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bf16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attn_tc -s 40 -c 2 -o gpurun_out/prof_attn_tc -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 80 -c 6 -o gpurun_out/prof_gemm_tc -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
