#!/bin/bash
# Round evidence pass: full parity suite, headline bench (fp32 with CPU baseline, bf16, reference arm), config 3 / 4 reports,
# geometry rows, ncu launch list of the bench command and full captures of the top kernels.  Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -c 300 gpurun_out/bench_fp32.err
timeout 600 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python tools/report_cfg4.py > gpurun_out/cfg4_report.json 2> gpurun_out/cfg4.err; cat gpurun_out/cfg4_report.json; tail -3 gpurun_out/cfg4.err
timeout 300 python tools/bench_window.py > gpurun_out/window_1gpu.json 2> gpurun_out/window.err; cat gpurun_out/window_1gpu.json
timeout 300 python tests/perf_geometry.py > gpurun_out/bench_geometry.jsonl 2> gpurun_out/bench_geometry.err
timeout 200 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1
timeout 200 python tools/time_stages.py fp32,bf16 > gpurun_out/time_stages.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 500 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attn_tc3 -s 6 -c 2 -o gpurun_out/prof_attn_tc3 -f python tools/prof_lg.py fp32 2 > gpurun_out/ncu_attn3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 20 -c 6 -o gpurun_out/prof_gemm_tc3 -f python tools/prof_lg.py fp32 2 > gpurun_out/ncu_gemm3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_preprocess|k_aliked_score|k_aliked_s8|k_sddh_sample|k_dkd_nms|k_conv3x3" -c 12 -o gpurun_out/prof_aliked -f python tools/prof_aliked.py 2 > gpurun_out/ncu_aliked.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_lg_lse2|k_lg_argmax2|k_lg_gather_blk|k_lg_heads_blk|k_ln_gelu|k_lg_posenc" -s 10 -c 12 -o gpurun_out/prof_lg_small -f python tools/prof_lg.py fp32 2 > gpurun_out/ncu_lgsmall.log 2>&1
timeout 300 ncu --set full --clock-control none -k "regex:k_fm_|k_remap|k_reproj" -c 12 -o gpurun_out/prof_geometry -f python tests/perf_geometry.py > gpurun_out/ncu_geometry.log 2>&1
# the .ncu-rep files are too large to travel back (64 MiB cap): keep their raw pages (and the attention kernel's source page) as CSV
for r in prof_attn_tc3 prof_gemm_tc3 prof_aliked prof_lg_small prof_geometry; do
  [ -f gpurun_out/$r.ncu-rep ] && ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
[ -f gpurun_out/prof_attn_tc3.ncu-rep ] && ncu -i gpurun_out/prof_attn_tc3.ncu-rep --page source --csv --print-source sass 2>/dev/null | head -c 3000000 > gpurun_out/prof_attn_tc3.source.csv
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out; ls -la gpurun_out
