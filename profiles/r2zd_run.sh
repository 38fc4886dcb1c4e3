#!/bin/bash
# GPU-box script of profiles/r2zd_* (2 GPUs): cost of the hook iteration on slabs (blocks of 1 ... 100 iterations)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 profiles/hook_cost.py 2>/dev/null | grep block | tee gpurun_out/r2zd_hook_cost_n2.txt
timeout 300 python profiles/hook_cost.py 2>/dev/null | grep block | tee gpurun_out/r2zd_hook_cost_n1.txt
