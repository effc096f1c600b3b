#!/bin/bash
# row-kernel probe, then what the driver runs at round end (gpu_full.sh)
mkdir -p gpurun_out
timeout -k 10 200 python tools/xl_probe.py --batch 128 --mode fused > gpurun_out/xl_probe.log 2>&1; head -2 gpurun_out/xl_probe.log
timeout -k 10 200 python tools/time_forward.py --batch 64 --reps 3 > gpurun_out/time_fwd.log 2>&1; tail -2 gpurun_out/time_fwd.log
T_ALL=400 T_BENCH=240 bash tools/gpu_full.sh
