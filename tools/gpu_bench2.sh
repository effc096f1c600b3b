#!/bin/bash
# 2-GPU bench (batch-sharded sampling + the training step with its NCCL all-reduce)
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench2 exit=$?"; tail -3 gpurun_out/bench2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ['value','n_gpus','ms_per_step','denoiser_step_ms']})
    print(json.dumps(d['configs'].get('train_step'),indent=0))
    print(json.dumps(d['configs'].get('px512_b16'),indent=0)[:600])
except Exception as e:
    print('parse error',e); print(open('gpurun_out/bench2.json').read()[:2000])
PY
