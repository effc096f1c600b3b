#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_ops_gpu.py -q --no-header -p no:cacheprovider -x -k "qkv_attention_fused" > gpurun_out/qa.log 2>&1
echo "exit=$?"; tail -5 gpurun_out/qa.log
timeout -k 5 240 python tools/qa_trace.py > gpurun_out/qa_trace.log 2>&1
echo "exit=$?"; head -48 gpurun_out/qa_trace.log
timeout -k 5 240 python tools/qa_probe.py --mode fused
