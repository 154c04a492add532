#!/bin/bash
# Round-2 evidence pass: ncu launch list of the bench command, full captures of the dominant kernels at batch 8, sanitizer.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/r2_launches_bench_fp32.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-window > gpurun_out/r2_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attn_tc3 -s 20 -c 2 -o gpurun_out/r2_prof_attn_tc3 -f python tools/prof_batch.py fp32 8 2 > gpurun_out/r2_ncu_attn3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tcp -s 30 -c 6 -o gpurun_out/r2_prof_gemm_tcp -f python tools/prof_batch.py fp32 8 2 > gpurun_out/r2_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:k_attn_tc$|k_gemm_tcp" -s 30 -c 7 -o gpurun_out/r2_prof_bf16 -f python tools/prof_batch.py bf16 8 2 > gpurun_out/r2_ncu_bf16.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:k_lg_|k_ln_gelu" -s 2 -c 14 -o gpurun_out/r2_prof_lg_small -f python tools/prof_batch.py fp32 8 1 > gpurun_out/r2_ncu_lgsmall.log 2>&1
for r in r2_prof_attn_tc3 r2_prof_gemm_tcp r2_prof_bf16 r2_prof_lg_small; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
rm -f gpurun_out/*.ncu-rep
bash tools/gpu_sanitize.sh r2 gemm attn match batch
du -sh gpurun_out; ls gpurun_out | head -50
