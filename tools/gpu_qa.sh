#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_forward_gpu.py tests/test_ops_gpu.py -q --no-header -p no:cacheprovider -x -k "not self_attention" 2>&1 | tail -4
timeout -k 5 300 python tools/time_forward.py --batch 64 --reps 3 | tail -2
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/l2.csv python tools/time_forward.py --batch 64 --reps 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/l2.csv | head -9
