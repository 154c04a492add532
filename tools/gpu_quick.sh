#!/bin/bash
# quick GPU pass: parity tests + both bench precisions (no cpu baseline)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/q_bf16.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bf16', d['value'], 'e2e', d['e2e']['value'], 'attn_ms', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'])"
B2S_NO_PDL=1 timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bf16 nopdl', d['value'], 'e2e', d['e2e']['value'])"
timeout 300 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/q_fp32.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fp32', d['value'], 'e2e', d['e2e']['value'])"
