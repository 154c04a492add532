#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_gpu.log
tail -4 gpurun_out/r2l_pytest_gpu.log
for b in 1 8; do for p in fp32 bf16; do timeout 120 python tools/prof_batch.py $p $b 4 1 2>/dev/null | tee -a gpurun_out/r2l_prof_classes.txt; done; done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l_bench_fp32.json 2> gpurun_out/r2l_bench_fp32.err; python -c "
import json;d=json.load(open('gpurun_out/r2l_bench_fp32.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches']});print('e2e',d['e2e']['value']);print('roof',d['roofline']['frac'],d['roofline']['avg_launch_ms'],'gemm',d['roofline_gemm']['frac'],d['roofline_gemm']['mma_issued_frac_of_peak']);print(d.get('other_precision',{}).get('value'));print(d.get('adaptive_leg',{}).get('value'));print(d.get('config3_window',{}).get('pairs_per_s'));print(d.get('parity_check'));print(d['clocks'])"
