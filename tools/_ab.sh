timeout 120 python tools/probe_h2.py 2>&1 | grep -E "gemm_h2|attn_h2|attn_tc "
timeout 200 python tools/time_batch.py fp32,bf16 1,8 2>&1 | grep -E "LightGlue|ALIKED"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
