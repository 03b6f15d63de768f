#!/bin/bash
# compute-sanitizer over the tcgen05 kernel tests and one small forward: memcheck, then racecheck.
# Run on a GPU box:  bash profiles/sanitize.sh   -> gpurun_out/r2_sanitizer_{memcheck,racecheck}.log
set -u
mkdir -p gpurun_out
SEL='linear_tc and (128-128-64 or 300-256-256 or 77-1536-256) or conv_tc and (32-32-9-5 or 128-34-32-18) or attention_tc or feed_forward and (34-1024 or 129-256) or layernorm_epilogue and 34-512 or fgd_statistics_match or logmel_tile_kernels_agree_bitwise and (1-True or 0-False) or logmel_matches_fp64_oracle and ted'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 \
    python -m pytest tests/test_gpu_tc_kernels.py tests/test_gpu_parity.py -m gpu -q -x -k "$SEL or ted_b2 and tc" \
    > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool exit code $?" >> gpurun_out/r2_sanitizer_$tool.log
  tail -4 gpurun_out/r2_sanitizer_$tool.log
done
