#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_vae.py tests/test_reference_suite_gpu.py -q --no-header -p no:cacheprovider -x 2>&1 | tail -8
timeout -k 5 300 python tools/vae_profile.py > gpurun_out/vae_profile.txt 2>&1
head -16 gpurun_out/vae_profile.txt | cut -c1-100,140-230 | grep -v "^-\|Warn\|_warn"
tail -3 gpurun_out/vae_profile.txt
