#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?"; tail -5 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print(json.dumps({k:d[k] for k in ['value','ms_per_step','denoiser_step_ms','denoiser_step_frac_of_sustained_bf16_peak','frac_of_sustained_bf16_peak_whole_step','e2e','clocks','gpu_launches']},indent=0))
    print(json.dumps(d['roofline'],indent=0)[:1500])
    print(json.dumps(d['configs'],indent=0)[:6000])
    print(json.dumps(d.get('stock_torch_b200'),indent=0)[:800])
    print(json.dumps(d.get('cpu_baseline'),indent=0)[:800])
except Exception as e:
    print('parse error',e); print(open('gpurun_out/bench.json').read()[:2000])
PY
