#!/bin/bash
# perf iteration pass: parity of everything that touches the GEMM / attention kernels, stage timings, GEMM phases, fp32 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_lightglue.py tests/test_gpu_aliked.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_perf.log
timeout 200 python tools/bench_gemm.py 2>&1 | tee gpurun_out/bench_gemm.txt
timeout 200 python tools/time_stages.py fp32 2>&1 | tee gpurun_out/time_stages.txt
timeout 300 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>gpurun_out/q.err | tail -1 | tee gpurun_out/q_fp32.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp32', d['value'], 'e2e', d['e2e']['value'], 'attn_ms', d['roofline']['avg_launch_ms'], 'gemm share', d['roofline']['gemm_share_of_match'], 'match ms', d['roofline']['match_ms_single_stream'])"; tail -3 gpurun_out/q.err
