#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step, one training step, + full sections of one decoder block's kernels
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 4400 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu-launches exit=$?" >> gpurun_out/summary.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1200 --csv --log-file gpurun_out/launches_train.csv \
  python tools/time_train.py --batch 32 --steps 2 > gpurun_out/train_under_ncu.log 2>&1
echo "ncu-train exit=$?" >> gpurun_out/summary.txt
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_bf16_tn|attention_tc|dwconv_gelu_g16|layernorm" -s 120 -c 9 -o gpurun_out/prof_block -f \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only --pdl 0 > gpurun_out/ncu_block.log 2>&1
echo "ncu-block exit=$?" >> gpurun_out/summary.txt
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi.csv
cat gpurun_out/summary.txt
