#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_all 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider
run time_fwd 300 python tools/time_forward.py --batch 64 --reps 3
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwd.csv \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_fwd.log 2>&1
timeout -k 10 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/time_fwd.log; cat gpurun_out/bench.json | cut -c1-600
