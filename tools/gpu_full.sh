#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-5} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_all 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
TAILN=2 run time_pdl0 300 python tools/time_forward.py --batch 64 --reps 3 --pdl 0
TAILN=2 run time_pdl1 300 python tools/time_forward.py --batch 64 --reps 3 --pdl 1
cat gpurun_out/summary.txt
