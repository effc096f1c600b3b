#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -12 gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_fwd 400 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py -q --no-header -p no:cacheprovider
run train32 200 python tools/time_train.py --batch 32
run train256 200 python tools/time_train.py --batch 256
run time256 120 python tools/time_forward.py --batch 64 --reps 3 --forward-only --pdl 0
cat gpurun_out/summary.txt
