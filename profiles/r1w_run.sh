#!/bin/bash
# GPU-box script of profiles/r1w_* (8 GPUs): C5 512^3 + DDI SIB, all-to-alls pipelined vs plain; 4-rank parity
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731"
timeout 400 $T profiles/bench_c5.py --edge 512 --steps 10 2>gpurun_out/r1w_err.txt | grep config | tee gpurun_out/r1w_bench_c5_512_n8.txt
SPIRIT_B200_DDI_PIPELINE=0 timeout 400 $T profiles/bench_c5.py --edge 512 --steps 10 2>>gpurun_out/r1w_err.txt | grep config | tee gpurun_out/r1w_bench_c5_512_n8_nopipe.txt
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "4" 2>&1 | tail -1 | tee gpurun_out/r1w_mgpu_n4.txt
