#!/bin/bash
# A/B on ONE box (box-to-box clock / power differences are larger than some of the effects measured)
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-6} gpurun_out/$name.log >> gpurun_out/summary.txt; }
TAILN=1 run time_emu6 300 python tools/time_forward.py --batch 64 --reps 3 --qkv-emu 6
TAILN=1 run time_emu0 300 python tools/time_forward.py --batch 64 --reps 3 --qkv-emu 0
TAILN=1 run time_emu4 300 python tools/time_forward.py --batch 64 --reps 3 --qkv-emu 4
TAILN=1 run time_emu8 300 python tools/time_forward.py --batch 64 --reps 3 --qkv-emu 8
TAILN=1 run time_emu6b 300 python tools/time_forward.py --batch 64 --reps 3 --qkv-emu 6
timeout -k 5 600 ncu --set full --import-source on --clock-control none -k regex:qkv_attention -s 3 -c 1 -o gpurun_out/qa_ncu -f python tools/qa_probe.py --mode fused --reps 2 > gpurun_out/qa_ncu.log 2>&1
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_vae.py -q --no-header -p no:cacheprovider -k "(conv_out3 and not 256-256) or (qkv_attention_fused and (2-128 or 5-192))" > gpurun_out/sanitize.log 2>&1
echo "sanitizer exit=$?" >> gpurun_out/summary.txt; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
