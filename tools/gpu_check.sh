#!/bin/bash
# One GPU-box pass for the round's evidence: parity tests, bench (fp32 headline incl. CPU baseline, bf16, reference arm),
# ncu launch list of the bench command and full captures of the top kernels.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -c 300 gpurun_out/bench_fp32.err
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attn_tc3 -s 6 -c 2 -o gpurun_out/prof_attn_tc3 -f python tools/prof_lg.py fp32 2 > gpurun_out/ncu_attn3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 20 -c 6 -o gpurun_out/prof_gemm_tc3 -f python tools/prof_lg.py fp32 2 > gpurun_out/ncu_gemm3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv3x3_tc -s 3 -c 3 -o gpurun_out/prof_conv_tc -f python tools/prof_aliked.py 2 > gpurun_out/ncu_conv.log 2>&1
ls -la gpurun_out
