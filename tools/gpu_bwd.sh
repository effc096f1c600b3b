#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_sab 240 python -m pytest tests/test_backward_gpu.py -q --no-header -p no:cacheprovider -k "self_attention" -x
if grep -q "t_sab exit=0" gpurun_out/summary.txt; then
run t_bwd 600 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py -q --no-header -p no:cacheprovider
TAILN=1 run train32 300 python tools/time_train.py --batch 32 --steps 5
TAILN=1 run train256 300 python tools/time_train.py --batch 256 --steps 3
fi
cat gpurun_out/summary.txt
