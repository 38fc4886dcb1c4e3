#!/bin/bash
# GPU-box script of profiles/r2j_* (2 GPUs): halo exchange by peer stores inside the fused kernel (CUDA IPC + stream memory
# operations) against the NCCL exchange: parity vs one GPU, weak-scaling bench
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tests/mgpu_worker.py > gpurun_out/r2j_mgpu_n2.txt 2>&1; echo "worker exit $?" | tee -a gpurun_out/r2j_mgpu_n2.txt
grep -E "OK|FAIL|MGPU|Error|error" gpurun_out/r2j_mgpu_n2.txt | cut -c1-170 | tail -45
for np in 0 1; do
SPIRIT_B200_NO_PEER=$np timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2j_bench_n2_nopeer$np.json 2> gpurun_out/r2j_bench_n2_nopeer$np.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2j_bench_n2_nopeer$np.json') if l.startswith('{')][-1]); print('NO_PEER=$np N=2: value %.4e ms/step %.4f e2e %.4e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
