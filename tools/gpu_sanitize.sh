#!/bin/bash
# compute-sanitizer over the tcgen05 / TMA / mbarrier kernels (SURVEY.md:298): memcheck, racecheck, synccheck, initcheck.
# usage: tools/gpu_sanitize.sh <tag> [targets...]
TAG=${1:-r2}; shift
TARGETS=${@:-gemm attn match}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for t in $TARGETS; do
    log=gpurun_out/${TAG}_sanitizer_${tool}_${t}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $t > $log 2>&1
    echo "rc=$?" >> $log
    echo "== $tool $t: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $log | tr '\n' ' ')"
  done
done
