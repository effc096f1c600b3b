#!/bin/bash
mkdir -p gpurun_out && rm -f gpurun_out/summary2.txt
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench2 exit=$?" >> gpurun_out/summary2.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench2_ref.json 2> gpurun_out/bench2_ref.err
echo "bench2 reference arm exit=$?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench2.json').read().strip().split('\n')[-1])
print({k:d.get(k) for k in ['value','n_gpus','ms_per_step','denoiser_step_ms','e2e','gpu_launches']})
c=d.get('configs',{})
print({k:(v.get('ms_per_step'),v.get('allreduce_exposed_ms')) for k,v in c.items() if isinstance(v,dict) and 'ms_per_step' in v})
print(c.get('px512_b16',{}).get('ms_per_model_call'))
PY
tail -c 600 gpurun_out/bench2_ref.json; tail -3 gpurun_out/bench2.err
