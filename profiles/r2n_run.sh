#!/bin/bash
# GPU-box script of profiles/r2n_* (2 GPUs): distributed dipolar convolution over peer-mapped memory (no all-to-all): parity vs one
# GPU, 256^3 + DDI SIB at N = 2 with and without the peer path
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tests/mgpu_worker.py > gpurun_out/r2n_mgpu_n2.txt 2>&1; echo "worker exit $?" | tee -a gpurun_out/r2n_mgpu_n2.txt
grep -E "DDI|FAIL|MGPU|Error|error" gpurun_out/r2n_mgpu_n2.txt | cut -c1-170 | tail -12
for np in 0 1; do
SPIRIT_B200_NO_PEER=$np timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c1-260
done
