#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_forward_gpu.py -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
for p in 1 0 1 0; do echo "share=$p"; timeout -k 5 300 python tools/time_forward.py --batch 64 --reps 3 --share-cfg $p | tail -2; done
