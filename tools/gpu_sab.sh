#!/bin/bash
mkdir -p gpurun_out
python tools/sab_probe.py 2>&1 | tail -8
