#!/bin/bash
mkdir -p gpurun_out
for p in 0 1; do
B2S_BENCH_PRIO=$p timeout 300 python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>gpurun_out/q.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prio $p fp32 value', d['value'], 'e2e', d['e2e']['value'])"
done
