#!/bin/bash
# GPU-box script of profiles/r2b_*: ncu --set full of the fused kernel (default build and variant fB)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_fused -s 6 -c 1 -o gpurun_out/r2b_prof_A -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_ncu_A.log 2>&1
tail -2 gpurun_out/r2b_ncu_A.log | cut -c1-200
SPIRIT_B200_LIB=libSpirit_fB.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_fused -s 6 -c 1 -o gpurun_out/r2b_prof_B -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2b_ncu_B.log 2>&1
tail -2 gpurun_out/r2b_ncu_B.log | cut -c1-200
ls -la gpurun_out/
