#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (small cases only)
mkdir -p gpurun_out
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_clip.py tests/test_backward_gpu.py tests/test_vae.py -q --no-header -p no:cacheprovider \
  -k "conv_out3 and not 256-256" > gpurun_out/sanitize.log 2>&1
echo "sanitizer exit=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize.log | head -20; tail -3 gpurun_out/sanitize.log
