#!/bin/bash
# round-2 evidence: full GPU suite, bench line, launch list of one sampler generation, ncu --set full of one decoder block
mkdir -p gpurun_out && rm -f gpurun_out/summary.txt
run() { name=$1; shift; timeout -k 10 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -${TAILN:-4} gpurun_out/$name.log >> gpurun_out/summary.txt; }
run t_all 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" >> gpurun_out/summary.txt
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 2400 --csv --log-file gpurun_out/launches_sampler.csv python tools/time_forward.py --batch 64 --reps 1 > gpurun_out/ncu_sampler.log 2>&1
python tools/launch_summary.py gpurun_out/launches_sampler.csv > gpurun_out/launch_summary_sampler.txt 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_bf16_tn|qkv_attention|gemm_up_dwconv|layernorm_bf16|ln_xattn_ln" -s 15 -c 5 -o gpurun_out/prof_block -f python tools/time_forward.py --batch 64 --reps 1 --forward-only > gpurun_out/ncu_block.log 2>&1
echo "ncu block exit=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; head -16 gpurun_out/launch_summary_sampler.txt
