#!/bin/bash
# First-light run on the B200 box: per-kernel parity, forward/sampler parity, quick timing, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in test_ops_gpu test_forward_gpu; do
  timeout -k 10 600 python -m pytest tests/$grp.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$grp.log 2>&1
  echo "$grp exit=$?" >> gpurun_out/summary.txt
  tail -5 gpurun_out/$grp.log >> gpurun_out/summary.txt
done
timeout -k 10 300 python tools/time_forward.py --batch 64 --reps 3 > gpurun_out/time_forward.log 2>&1
echo "time_forward exit=$?" >> gpurun_out/summary.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwd.csv \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_fwd.log 2>&1
echo "ncu exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/time_forward.log
