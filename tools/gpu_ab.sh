#!/bin/bash
# A/B on ONE box: fused qkv + attention on/off (box-to-box clock / power differences are larger than some of the effects measured)
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
TAILN=3 run t_qa 300 python -m pytest tests/test_ops_gpu.py -q --no-header -p no:cacheprovider -x -k "qkv_attention_fused"
TAILN=2 run time_qkv1 300 python tools/time_forward.py --batch 64 --reps 3 --fused-qkv 1
TAILN=2 run time_qkv0 300 python tools/time_forward.py --batch 64 --reps 3 --fused-qkv 0
TAILN=2 run time_qkv1b 300 python tools/time_forward.py --batch 64 --reps 3 --fused-qkv 1
TAILN=40 run trace 200 python tools/qa_trace.py
TAILN=2 run probe 200 python tools/qa_probe.py --mode fused
TAILN=8 run t_fwd 600 python -m pytest tests/test_forward_gpu.py -q --no-header -p no:cacheprovider -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
