#!/bin/bash
# GPU-box script of profiles/r2e_*: plain march (rotated loop, role split, no branches inside the gradient) against the general march
mkdir -p gpurun_out
SPIRIT_B200_LIB=libSpirit.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "fused or iterate_block or single_steps or fullsize or 256" > gpurun_out/r2e_pytest.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2e_pytest.txt
tail -3 gpurun_out/r2e_pytest.txt
timeout 900 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit_fM.so" "LIB=libSpirit_fN.so" "LIB=libSpirit_fL.so" "LIB=libSpirit.so" "LIB=libSpirit.so FUSED_LC=64" "LIB=libSpirit.so FUSED_LC=22" > gpurun_out/r2e_sweep.txt 2>&1
cat gpurun_out/r2e_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_fused -s 6 -c 1 -o gpurun_out/r2e_prof_main -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_ncu_main.log 2>&1
