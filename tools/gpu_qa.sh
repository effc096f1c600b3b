#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests/test_vae.py -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout -k 5 300 python tools/vae_profile.py > gpurun_out/vae_profile.txt 2>&1
head -12 gpurun_out/vae_profile.txt | cut -c1-100,140-230 | grep -v "^-\|Warn\|_warn"
tail -3 gpurun_out/vae_profile.txt
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --headline-only > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench exit=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_h.json').read().strip().split('\n')[-1]); print({k:d[k] for k in ['value','ms_per_step','denoiser_step_ms','e2e','clocks']})"
