#!/bin/bash
# GPU-box script of profiles/r2zm_* (2 GPUs): the multi-GPU worker and the bench line at N = 2 with the register-resident dipolar passes
# (k_fft_pass16r reads / writes the peer-mapped operands of the pencil decomposition)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29716 tests/mgpu_worker.py > gpurun_out/r2zm_mgpu_n2.txt 2>&1; echo "worker exit $?" | tee -a gpurun_out/r2zm_mgpu_n2.txt
grep -E "OSO|Atlas|FAIL|MGPU|Error|error" gpurun_out/r2zm_mgpu_n2.txt | cut -c1-200 | tail -30
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2zm_bench_n2.json 2> gpurun_out/r2zm_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2zm_bench_n2.json') if l.startswith('{')][-1])
print('ms/step %.4f' % d['ms_per_step'], 'e2e', d['e2e']['value'])
print('c4', d['configs'].get('c4'))
print('c5', d['configs'].get('c5'))
print('parity', d['multi_gpu_parity'])
PY
tail -5 gpurun_out/r2zm_bench_n2.err
