#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/tests_gpu.log 2>&1
echo "tests exit=$?" >> gpurun_out/summary.txt; tail -4 gpurun_out/tests_gpu.log >> gpurun_out/summary.txt
timeout -k 10 300 python tools/time_forward.py --batch 64 --reps 3 > gpurun_out/time_forward.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwd.csv \
  python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_fwd.log 2>&1
timeout -k 10 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/time_forward.log gpurun_out/bench.json; tail -3 gpurun_out/bench.err gpurun_out/smoke.log
