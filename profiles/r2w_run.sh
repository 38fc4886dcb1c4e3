#!/bin/bash
# GPU-box script of profiles/r2w_* (2 GPUs): minimisers (VP_OSO / LBFGS_OSO / LBFGS_Atlas) on slabs and on sharded chains against one
# GPU; the whole multi-GPU worker; bench line at N = 2 (c4 on a sharded chain, c5, parity record)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tests/mgpu_worker.py > gpurun_out/r2w_mgpu_n2.txt 2>&1; echo "worker exit $?" | tee -a gpurun_out/r2w_mgpu_n2.txt
grep -E "OSO|Atlas|FAIL|MGPU|Error|error" gpurun_out/r2w_mgpu_n2.txt | cut -c1-200 | tail -30
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2w_bench_n2.json') if l.startswith('{')][-1])
print('ms/step %.4f' % d['ms_per_step'], 'e2e', d['e2e']['value'])
print('c4', d['configs'].get('c4'))
print('c5', d['configs'].get('c5'))
print('parity', d['multi_gpu_parity'])
PY
tail -5 gpurun_out/r2w_bench_n2.err
