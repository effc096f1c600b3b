#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
for impl in 2 3; do
python tools/attn_probe.py --impl $impl >> gpurun_out/summary.txt 2>&1
python tools/attn_probe.py --impl $impl --batch 8 --tok 4096 >> gpurun_out/summary.txt 2>&1
done
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:attention_tc2 -s 2 -c 1 -o gpurun_out/attn_tc2_4096 -f \
  python tools/attn_probe.py --impl 3 --batch 8 --tok 4096 --reps 1 > gpurun_out/ncu_a.log 2>&1
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:attention_tc2 -s 2 -c 1 -o gpurun_out/attn_tc2_256 -f \
  python tools/attn_probe.py --impl 3 --reps 1 > gpurun_out/ncu_b.log 2>&1
cat gpurun_out/summary.txt
