#!/bin/bash
# fused up-projection + dwconv + GELU kernel: parity first (short timeout: a barrier bug would hang), then timings
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_fused 240 python -m pytest tests/test_ops_gpu.py -q -k "fused" --no-header -p no:cacheprovider -x
if grep -q "t_fused exit=0" gpurun_out/summary.txt; then
  run t_fwd 600 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py -q --no-header -p no:cacheprovider -x
  TAILN=8 run time_fused 300 python tools/time_forward.py --batch 64 --reps 5 --fused-mlp 1
  TAILN=8 run time_unfused 300 python tools/time_forward.py --batch 64 --reps 5 --fused-mlp 0
fi
cat gpurun_out/summary.txt
