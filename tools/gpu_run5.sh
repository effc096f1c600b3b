#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_bf16_tn" -s 60 -c 4 -o gpurun_out/prof_gemm2 -f \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only --gemm-ctas 2 > gpurun_out/ncu_gemm2.log 2>&1
echo "ncu-gemm2 exit=$?" >> gpurun_out/summary.txt
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:attention_tc|dwconv" -s 20 -c 2 -o gpurun_out/prof_attn -f \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only --gemm-ctas 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu-attn exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
