timeout 600 python -m pytest tests/test_gpu_aliked.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/time_extract_batch.py 2>&1 | grep "lanes="
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2am_launches_aliked.csv python tools/prof_aliked.py 3 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2am_launches_aliked.csv | head -14
