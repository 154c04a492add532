#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_gpu.log
tail -12 gpurun_out/r2e_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_fp32.json 2> gpurun_out/r2e_bench_fp32.err; tail -c 600 gpurun_out/r2e_bench_fp32.err; python -c "
import json;d=json.load(open('gpurun_out/r2e_bench_fp32.json'));print({k:d[k] for k in ['value','ms_per_step','gpu_launches']});print('e2e',d['e2e']);print('roof',d['roofline']['frac'],d['roofline']['avg_launch_ms'],'gemm',d['roofline_gemm']['frac'],d['roofline_gemm']['mma_issued_frac_of_peak']);print(d.get('other_precision'));print(d.get('adaptive_leg'));print(d.get('config3_window'));print(d.get('parity_check'));print(d.get('cpu_baseline'));print(d['clocks'])"
for l in 1 2 4 8; do timeout 300 python bench.py --steps 6 --warmup 2 --lanes $l --no-cpu-baseline --no-secondary --no-window 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('lanes $l value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"; done
for b in 2 4 16; do timeout 300 python bench.py --steps 6 --warmup 2 --batch $b --no-cpu-baseline --no-secondary --no-window 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('batch $b value',round(d['value'],1))"; done
