#!/bin/bash
# A/B on ONE box: fused MLP on/off (box-to-box clock / power differences are larger than some of the effects measured)
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
TAILN=2 run time_fused1 300 python tools/time_forward.py --batch 64 --reps 3 --fused-mlp 1
TAILN=2 run time_fused0 300 python tools/time_forward.py --batch 64 --reps 3 --fused-mlp 0
TAILN=2 run time_fused1b 300 python tools/time_forward.py --batch 64 --reps 3 --fused-mlp 1
run t_adam 300 python -m pytest tests/test_train_gpu.py -q --no-header -p no:cacheprovider -k adam
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
