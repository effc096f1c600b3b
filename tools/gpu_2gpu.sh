#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary2.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench2 exit=$?" >> gpurun_out/summary2.txt
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/time_train.py --batch 32 > gpurun_out/train2.log 2>&1
echo "train2 exit=$?" >> gpurun_out/summary2.txt
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/time_train.py --batch 128 > gpurun_out/train2b.log 2>&1
echo "train2b exit=$?" >> gpurun_out/summary2.txt
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 tools/time_train.py --batch 32 --no-overlap > gpurun_out/train2c.log 2>&1
echo "train2c exit=$?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt; tail -c 900 gpurun_out/bench2.json | head -c 900; echo; tail -2 gpurun_out/train2.log; tail -2 gpurun_out/train2b.log; tail -2 gpurun_out/train2c.log; tail -3 gpurun_out/bench2.err
