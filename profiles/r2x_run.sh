#!/bin/bash
# GPU-box script of profiles/r2x_* (2 GPUs): dipolar convolution with ka pencils (transposes = stores / loads of the a-pass kernels,
# half the volume of the kb-block transposes): parity against one GPU, 512^3 + DDI SIB against the kb-block path
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tests/mgpu_worker.py > gpurun_out/r2x_mgpu_n2.txt 2>&1; echo "worker exit $?" | tee -a gpurun_out/r2x_mgpu_n2.txt
grep -E "DDI|FAIL|MGPU|Error|error" gpurun_out/r2x_mgpu_n2.txt | cut -c1-200 | tail -20
for P in 1 0; do
echo "== SPIRIT_B200_DDI_PENCIL=$P" | tee -a gpurun_out/r2x_sweep.txt
SPIRIT_B200_DDI_PENCIL=$P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 profiles/bench_c5.py --edge 512 --steps 5 2>gpurun_out/r2x_err_$P.txt | grep config | cut -c1-260 | tee -a gpurun_out/r2x_sweep.txt
tail -3 gpurun_out/r2x_err_$P.txt | cut -c1-300
done
timeout 300 python profiles/bench_c5.py --edge 512 --steps 5 2>/dev/null | grep config | cut -c1-260 | tee -a gpurun_out/r2x_sweep.txt
