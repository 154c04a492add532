#!/bin/bash
# Round-2 final evidence pass (one gpurun call): full GPU test log, bench lines (ours fp32 / bf16, reference arm), ncu launch
# list of the bench command, full captures (raw + source pages) of the dominant kernels at 8 pairs per launch, extractor
# launch list, sanitizer over the tensor-core entry points / match / batch / geometry.
set -x
mkdir -p gpurun_out
T=r2g
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt
python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${T}_pytest_gpu.log
python bench.py > gpurun_out/${T}_bench_fp32.json 2> gpurun_out/${T}_bench_fp32.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_cpu.json 2> gpurun_out/${T}_bench_reference_cpu.err
python bench.py --precision bf16 --no-cpu-baseline --no-window > gpurun_out/${T}_bench_bf16.json 2> gpurun_out/${T}_bench_bf16.err
python tools/time_batch.py fp32,bf16 1,8 2>&1 | grep -E "LightGlue|ALIKED" > gpurun_out/${T}_time_batch.txt
python tools/time_extract_batch.py > gpurun_out/${T}_time_extract_batch.txt 2>&1
python tools/prof_batch.py fp32 8 3 1 2>&1 | tail -1 > gpurun_out/${T}_batch_classes.txt
python tools/probe_h2.py 2>&1 | grep -E "gemm|attn" > gpurun_out/${T}_probe_h2.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/${T}_launches_bench_fp32.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-window > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_aliked.csv python tools/prof_aliked.py 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_attn_tc3|k_gemm_tcp|k_ln_gelu" -s 9 -c 10 -o gpurun_out/${T}_prof_layer -f python tools/prof_batch.py fp32 8 1 > gpurun_out/${T}_ncu_layer.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:k_attn_tc$|k_gemm_tcp" -s 30 -c 7 -o gpurun_out/${T}_prof_bf16 -f python tools/prof_batch.py bf16 8 2 > gpurun_out/${T}_ncu_bf16.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:k_lg_|k_fmcv" -s 2 -c 14 -o gpurun_out/${T}_prof_lg_small -f python tools/prof_batch.py fp32 8 1 > gpurun_out/${T}_ncu_lgsmall.log 2>&1
for r in ${T}_prof_layer ${T}_prof_bf16 ${T}_prof_lg_small; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
[ -f gpurun_out/${T}_prof_layer.ncu-rep ] && ncu -i gpurun_out/${T}_prof_layer.ncu-rep --page source --csv > gpurun_out/${T}_prof_layer.source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
bash tools/gpu_sanitize.sh ${T} gemm attn match batch geom
du -sh gpurun_out; ls gpurun_out | head -60
