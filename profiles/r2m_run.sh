#!/bin/bash
# GPU-box script of profiles/r2m_* (8 GPUs): the driver's scaling command at N = 8: weak scaling of the stencil step, parity record,
# strong scaling of configs[4] (512^3 + dipolar convolution)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n8.json 2> gpurun_out/r2m_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2m_bench_n8.json') if l.startswith('{')][-1])
print('N=8 value %.4e ms/step %.4f e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print(json.dumps(d['multi_gpu_parity'], indent=1))
print(json.dumps(d['configs'], indent=1))
PY
tail -5 gpurun_out/r2m_bench_n8.err
