#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 1000 --csv --log-file gpurun_out/launches_train.csv \
  python tools/time_train.py --batch 32 --steps 2 > gpurun_out/train_under_ncu.log 2>&1
echo "exit=$?"
