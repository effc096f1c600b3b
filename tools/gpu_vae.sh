#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_vae 900 python -m pytest tests/test_vae.py tests/test_reference_suite_gpu.py -q --no-header -p no:cacheprovider -s
TAILN=4 run time_vae 300 python tools/time_vae.py
grep -h "full-width" gpurun_out/t_vae.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
