#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_fused 240 python -m pytest tests/test_ops_gpu.py -q -k "fused" --no-header -p no:cacheprovider -x
run probe 240 python tools/fused_probe.py
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_up_dwconv -s 3 -c 1 -o gpurun_out/prof_fused -f python tools/fused_probe.py --mode fused --reps 2 > gpurun_out/ncu_fused.log 2>&1
echo "ncu exit=$?" >> gpurun_out/summary.txt
run t_adam 300 python -m pytest tests/test_train_gpu.py -q --no-header -p no:cacheprovider -x -k "adam"
cat gpurun_out/summary.txt
