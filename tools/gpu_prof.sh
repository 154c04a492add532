#!/bin/bash
# stage timings + ncu launch list + full captures of the top kernels (bf16 path)
mkdir -p gpurun_out
timeout 600 python tools/time_stages.py 2>&1 | tee gpurun_out/time_stages.txt
timeout 300 python tools/bench_kernels.py attn 2>&1 | tee gpurun_out/bench_attn.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 500 --csv --log-file gpurun_out/launches_bf16.csv python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bf16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attn_tc -s 40 -c 2 -o gpurun_out/prof_attn_tc -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 80 -c 6 -o gpurun_out/prof_gemm_tc -f python bench.py --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
