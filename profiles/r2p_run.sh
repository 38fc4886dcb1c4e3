#!/bin/bash
# GPU-box script of profiles/r2p_*: shared-memory slot index pinned in a register; hook iteration with the energy taken right after
# the gradient
mkdir -p gpurun_out
timeout 900 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit_fV.so" "LIB=libSpirit.so" "LIB=libSpirit_fV.so" > gpurun_out/r2p_sweep.txt 2>&1
cat gpurun_out/r2p_sweep.txt
python profiles/hook_cost.py 2>/dev/null | grep block | tee gpurun_out/r2p_hook_cost.txt
SPIRIT_B200_LIB=libSpirit_fV.so python profiles/hook_cost.py 2>/dev/null | grep block | tee gpurun_out/r2p_hook_cost_fV.txt
