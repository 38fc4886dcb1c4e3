#!/bin/bash
# GPU-box script of profiles/r1v_*: C5 (512^3, exchange + DMI + DDI, SIB) on ONE GPU: the denominator of the 8-GPU strong-scaling figure
mkdir -p gpurun_out
timeout 600 python profiles/bench_c5.py --edge 512 --steps 5 2>gpurun_out/r1v_err.txt | grep config | tee gpurun_out/r1v_bench_c5_512_n1.txt
tail -3 gpurun_out/r1v_err.txt
nvidia-smi --query-gpu=memory.used,memory.total --format=csv | tail -1
