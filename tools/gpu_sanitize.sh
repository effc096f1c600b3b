#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (small cases only)
mkdir -p gpurun_out
timeout -k 10 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_vae.py tests/test_clip.py -q --no-header -p no:cacheprovider \
  -k "(conv_out3 and not 256-256) or (qkv_attention_fused and (2-128 or 5-192)) or (ln_xattn_ln_fused and (2-64-128 or 3-36-512 or 5-32-256)) or (conv3x3_fused and 2-16-16) or (vae_attention_core and 64) or (fused and 64-256) or (encode_text and 128)" > gpurun_out/sanitize.log 2>&1
echo "sanitizer exit=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize.log | head -20; tail -3 gpurun_out/sanitize.log
