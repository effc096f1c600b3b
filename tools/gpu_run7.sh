#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_vae 600 python -m pytest tests/test_vae.py -q -m gpu --no-header -p no:cacheprovider
run time_vae 600 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from transformer_latent_diffusion_b200.vae import AutoencoderKLDecoder
for chunk in (16, 32, 64):
    m = AutoencoderKLDecoder(chunk=chunk).cuda().to(torch.bfloat16).eval().to(memory_format=torch.channels_last)
    z = torch.randn(64, 4, 32, 32, device="cuda")
    for _ in range(2): m.decode(z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(3): m.decode(z)
    e1.record(); torch.cuda.synchronize()
    print(f"chunk={chunk}: {e0.elapsed_time(e1)/3:.1f} ms per 64 images", flush=True)
from torch.profiler import profile, ProfilerActivity
m = AutoencoderKLDecoder(chunk=16).cuda().to(torch.bfloat16).eval().to(memory_format=torch.channels_last)
z = torch.randn(16, 4, 32, 32, device="cuda")
m.decode(z); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.decode(z); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
PY
timeout -k 10 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/time_vae.log | cut -c1-200; cat gpurun_out/bench.json | cut -c1-1200
