#!/bin/bash
# full GPU regression + stage timings + fp32 bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 200 python tools/time_stages.py fp32,bf16 2>&1 | tee gpurun_out/time_stages.txt
B2S_NO_PDL=1 timeout 200 python tools/time_stages.py fp32 2>&1 | sed 's/^/nopdl: /' | tee -a gpurun_out/time_stages.txt
timeout 300 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>gpurun_out/q.err | tail -1 | tee gpurun_out/q_fp32.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp32', d['value'], 'e2e', d['e2e']['value'], 'attn_ms', d['roofline']['avg_launch_ms'], 'gemm share', d['roofline']['gemm_share_of_match'], 'match ms', d['roofline']['match_ms_single_stream'])"; tail -3 gpurun_out/q.err
