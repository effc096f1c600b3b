#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_attn 120 python -m pytest tests/test_ops_gpu.py -q -k attention --no-header -p no:cacheprovider
run t_fwd 400 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py tests/test_reference_suite_gpu.py -q --no-header -p no:cacheprovider
for emu in 4 6 8; do
timeout -k 5 60 python tools/attn_probe.py --impl 3 --exp-emu $emu >> gpurun_out/summary.txt 2>&1
timeout -k 5 60 python tools/attn_probe.py --impl 3 --exp-emu $emu --batch 8 --tok 4096 >> gpurun_out/summary.txt 2>&1
done
timeout -k 5 60 python tools/attn_probe.py --impl 3 --batch 32 --tok 1024 >> gpurun_out/summary.txt 2>&1
run time256 120 python tools/time_forward.py --batch 64 --reps 3 --forward-only
run time512 120 python tools/time_forward.py --img 64 --batch 16 --reps 3 --forward-only
run time1024 120 python tools/time_forward.py --img 128 --batch 4 --reps 3 --forward-only
timeout -k 10 300 ncu --set full --import-source on --clock-control none -k regex:attention_tc2 -s 2 -c 1 -o gpurun_out/attn_tc2_256 -f \
  python tools/attn_probe.py --impl 3 --reps 1 > gpurun_out/ncu_b.log 2>&1
cat gpurun_out/summary.txt
