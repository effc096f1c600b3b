#!/bin/bash
# backward-kernel parity + training-step timing (one GPU)
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_bwd 400 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py -q --no-header -p no:cacheprovider -x
run train32 200 python tools/time_train.py --batch 32
run train256 200 python tools/time_train.py --batch 256
cat gpurun_out/summary.txt
