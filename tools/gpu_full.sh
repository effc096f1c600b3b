#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -4 gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_all 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/bench.json | cut -c1-3000
