#!/bin/bash
# parity + step timings + per-kernel launch list
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_ops 600 python -m pytest tests/test_ops_gpu.py -q --no-header -p no:cacheprovider -x -k "gemm or fold or producer"
run t_fwd 900 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py -q --no-header -p no:cacheprovider
TAILN=3 run time_a 300 python tools/time_forward.py --batch 64 --reps 4
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 90 -c 190 --csv --log-file gpurun_out/launches_fwd.csv python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_fwd.log 2>&1
python tools/launch_summary.py gpurun_out/launches_fwd.csv > gpurun_out/launch_summary.txt 2>&1
cat gpurun_out/summary.txt; head -14 gpurun_out/launch_summary.txt
