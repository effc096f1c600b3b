#!/bin/bash
# Parity + timing + ncu (launch list and full sections of the GEMM / dwconv / attention kernels).
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
for grp in test_ops_gpu test_forward_gpu; do
  timeout -k 10 600 python -m pytest tests/$grp.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$grp.log 2>&1
  echo "$grp exit=$?" >> gpurun_out/summary.txt
  tail -4 gpurun_out/$grp.log >> gpurun_out/summary.txt
done
timeout -k 10 300 python tools/time_forward.py --batch 64 --reps 3 > gpurun_out/time_forward.log 2>&1
echo "time_forward exit=$?" >> gpurun_out/summary.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwd.csv \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_fwd.log 2>&1
echo "ncu-list exit=$?" >> gpurun_out/summary.txt
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 60 -c 4 -o gpurun_out/prof_gemm -f \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_gemm.log 2>&1
echo "ncu-gemm exit=$?" >> gpurun_out/summary.txt
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:dwconv|self_attention|layernorm" -s 30 -c 5 -o gpurun_out/prof_row -f \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_row.log 2>&1
echo "ncu-row exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/time_forward.log
