#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-8} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_clip 600 python -m pytest tests/test_clip.py -q --no-header -p no:cacheprovider
cat gpurun_out/summary.txt
