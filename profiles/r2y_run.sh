#!/bin/bash
# GPU-box script of profiles/r2y_* (8 GPUs): 512^3 + DDI SIB with ka pencils, phase timing; the bench line at N = 8
mkdir -p gpurun_out
SPIRIT_B200_DDI_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 profiles/bench_c5.py --edge 512 --steps 8 2>gpurun_out/r2y_c5_timing.err | grep config | cut -c1-260 | tee gpurun_out/r2y_c5_n8.txt
grep "ddi timing rank [07]:" gpurun_out/r2y_c5_timing.err | tail -4 | tee -a gpurun_out/r2y_c5_n8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 profiles/bench_c5.py --edge 512 --steps 8 2>/dev/null | grep config | cut -c1-260 | tee -a gpurun_out/r2y_c5_n8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2y_bench_n8.json') if l.startswith('{')][-1])
print('ms/step %.4f value %.4g' % (d['ms_per_step'], d['value']), 'e2e', d['e2e']['value'])
print('c4', d['configs'].get('c4'))
print('c5', d['configs'].get('c5'))
print('parity', d['multi_gpu_parity'])
PY
tail -3 gpurun_out/r2y_bench_n8.err | cut -c1-300
