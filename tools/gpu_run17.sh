#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_ops 300 python -m pytest tests/test_ops_gpu.py -q --no-header -p no:cacheprovider
run t_fwd 400 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py tests/test_reference_suite_gpu.py -q --no-header -p no:cacheprovider
run time256 120 python tools/time_forward.py --batch 64 --reps 3 --forward-only
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwd.csv \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_fwd.log 2>&1
cat gpurun_out/summary.txt
