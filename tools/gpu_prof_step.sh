#!/bin/bash
# launch list of one sampler step region (shares) + per-kernel metrics of one decoder block
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 90 -c 190 --csv --log-file gpurun_out/launches_fwd.csv python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_fwd.log 2>&1
echo "ncu exit=$?"
python tools/launch_summary.py gpurun_out/launches_fwd.csv 2>&1 | head -30
