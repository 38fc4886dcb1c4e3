#!/bin/bash
# GPU-box script of profiles/r1f_*: parity tests of the lean march, A/B sweep against the r1e build, ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1f_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1f_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1f_pytest.txt
tail -3 gpurun_out/r1f_pytest.txt
timeout 900 python profiles/sweep.py "LIB=libSpirit_base.so" "LIB=libSpirit.so" "LIB=libSpirit_v2w384.so" "LIB=libSpirit_v2w384.so BX=64" "LIB=libSpirit_v1w320.so BX=64" "LIB=libSpirit.so LC=16" "LIB=libSpirit.so LC=64" "LIB=libSpirit_base.so" "LIB=libSpirit.so" > gpurun_out/r1f_sweep.txt 2>&1
cat gpurun_out/r1f_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_stage -s 60 -c 2 -o gpurun_out/r1f_prof -f python bench.py --steps 5 --warmup 30 --no-e2e --no-cpu-baseline > gpurun_out/r1f_ncu.log 2>&1
tail -2 gpurun_out/r1f_ncu.log
