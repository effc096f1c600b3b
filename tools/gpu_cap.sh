#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 200 ncu --set full --clock-control none --import-source on -k "regex:ln_xattn_ln" -s 3 -c 1 -o gpurun_out/prof_xl_final -f python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_xl_final.log 2>&1
echo "ncu exit=$?"
