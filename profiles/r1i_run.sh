#!/bin/bash
# GPU-box script of profiles/r1i_*: stored-noise variant of the marching kernels: tests, A/B, bench, ncu.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1i_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1i_pytest.txt
tail -4 gpurun_out/r1i_pytest.txt
timeout 900 python profiles/sweep.py "LIB=libSpirit_noxi.so" "LIB=libSpirit.so" "LIB=libSpirit_noxi.so" "LIB=libSpirit.so" "LIB=libSpirit.so LC=16" "LIB=libSpirit.so LC=64" > gpurun_out/r1i_sweep.txt 2>&1
cat gpurun_out/r1i_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6 -s 60 -c 2 -o gpurun_out/r1i_prof -f python bench.py --steps 5 --warmup 30 --no-e2e --no-cpu-baseline > gpurun_out/r1i_ncu.log 2>&1
tail -2 gpurun_out/r1i_ncu.log
