#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_fused 240 python -m pytest tests/test_ops_gpu.py -q -k "fused" --no-header -p no:cacheprovider -x
run probe 240 python tools/fused_probe.py
run t_fwd 600 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py -q --no-header -p no:cacheprovider
TAILN=3 run time_a 300 python tools/time_forward.py --batch 64 --reps 4
cat gpurun_out/summary.txt
