#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn_kernel -s 20 -c 8 -o gpurun_out/prof_lnp -f python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_lnp.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/ncu_lnp.log
