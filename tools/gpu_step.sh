#!/bin/bash
# forward/sampler parity + step timings (fused MLP on / off)
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_fwd 900 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py tests/test_ops_gpu.py -q --no-header -p no:cacheprovider -x
TAILN=4 run time_a 300 python tools/time_forward.py --batch 64 --reps 5 --fused-mlp 1
TAILN=4 run time_b 300 python tools/time_forward.py --batch 64 --reps 5 --fused-mlp 0
cat gpurun_out/summary.txt
