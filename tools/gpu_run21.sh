#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -12 gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_new 600 python -m pytest tests/test_vae.py tests/test_reference_suite_gpu.py -q -m gpu --no-header -p no:cacheprovider -x
cat gpurun_out/summary.txt
