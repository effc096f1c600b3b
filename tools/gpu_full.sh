#!/bin/bash
# what the driver runs at round end: pytest -m gpu, __graft_entry__.smoke(), bench.py
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-5} gpurun_out/$name.log >> gpurun_out/summary.txt; }
TAILN=16 run t_all ${T_ALL:-1500} python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --durations=10
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout -k 10 ${T_BENCH:-900} python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; head -c 1500 gpurun_out/bench.json
