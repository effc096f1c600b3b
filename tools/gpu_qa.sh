#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_ops_gpu.py -q --no-header -p no:cacheprovider -x -k "ln_xattn_ln_fused" 2>&1 | tail -5
timeout -k 5 240 python tools/xl_probe.py
timeout -k 5 300 python tools/time_forward.py --batch 64 --reps 3 | tail -2
timeout -k 5 600 ncu --set full --import-source on --clock-control none -k regex:ln_xattn_ln -s 3 -c 1 -o gpurun_out/xl_ncu -f python tools/xl_probe.py --mode fused --reps 2 > gpurun_out/xl_ncu.log 2>&1
echo "ncu exit=$?"
