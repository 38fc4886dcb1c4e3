#!/bin/bash
# GPU-box script of profiles/r1r_* (2 GPUs): distributed dipole convolution with the all-to-alls pipelined per component
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "2" > gpurun_out/r1r_mgpu_n2.txt 2>&1; grep -i "MGPU\|passed\|failed\|Error" gpurun_out/r1r_mgpu_n2.txt | tail -5
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721"
timeout 300 $T profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | tee gpurun_out/r1r_bench_c5_256_n2.txt
SPIRIT_B200_DDI_PIPELINE=0 timeout 300 $T profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | tee gpurun_out/r1r_bench_c5_256_n2_nopipe.txt
timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | tee gpurun_out/r1r_bench_c5_256_n1.txt
