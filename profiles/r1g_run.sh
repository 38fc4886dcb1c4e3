#!/bin/bash
# GPU-box script of profiles/r1g_*: parity of the tile-staged march, A/B sweep, ncu captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tile_gpu.py -q > gpurun_out/r1g_pytest_tile.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1g_pytest_tile.txt
grep -v "^E  \|^  \|^$\|^2026\|=====\|^tests/\|^cfg\|^product\|^oracle\|^monkey\|^solver\|^extra\|^    " gpurun_out/r1g_pytest_tile.txt | tail -25
timeout 900 python profiles/sweep.py "LIB=libSpirit_base.so" "LIB=libSpirit.so" "LIB=libSpirit.so SPIRIT_B200_SC6_TILED=0" "LIB=libSpirit.so SPIRIT_B200_SC6T_NS=2" "LIB=libSpirit.so LC=32" "LIB=libSpirit.so LC=16" "LIB=libSpirit.so SPIRIT_B200_SC6_TILED=0 LC=32" > gpurun_out/r1g_sweep.txt 2>&1
cat gpurun_out/r1g_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6 -s 60 -c 2 -o gpurun_out/r1g_prof -f python bench.py --steps 5 --warmup 30 --no-e2e --no-cpu-baseline > gpurun_out/r1g_ncu.log 2>&1
tail -2 gpurun_out/r1g_ncu.log
