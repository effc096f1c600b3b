#!/bin/bash
# norm2 + cross-attention + norm3 row kernels (FFMA rows per warp x CTAs per SM, tensor-pipe variants): parity, the kernel alone, the sampler, ncu capture
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-4} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_xm 400 python -m pytest tests/test_ops_gpu.py tests/test_forward_gpu.py -q -m gpu --no-header -p no:cacheprovider -k "ln_xattn_ln or fused_attention_kernels"
TAILN=10 run xl_probe 200 python tools/xl_probe.py --batch 128 --mode fused
for r in 4 2; do TAILN=2 run time_rows$r 200 python tools/time_forward.py --batch 64 --reps 3 --xattn-rows $r; done
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k "regex:ln_xattn_ln" -s 3 -c 1 -o gpurun_out/prof_xl_hb -f python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_xl_hb.log 2>&1
echo "ncu exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
