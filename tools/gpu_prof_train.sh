#!/bin/bash
mkdir -p gpurun_out
python tools/time_train.py --batch 32 --steps 5 > gpurun_out/time_train32.log 2>&1; tail -1 gpurun_out/time_train32.log
python tools/time_train.py --batch 256 --steps 3 > gpurun_out/time_train256.log 2>&1; tail -1 gpurun_out/time_train256.log
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1400 -c 700 --csv --log-file gpurun_out/launches_train.csv python tools/time_train.py --batch 32 --steps 2 > gpurun_out/ncu_train.log 2>&1
python tools/launch_summary.py gpurun_out/launches_train.csv > gpurun_out/launch_summary_train.txt 2>&1; head -40 gpurun_out/launch_summary_train.txt
