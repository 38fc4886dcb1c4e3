#!/bin/bash
# GPU-box script of profiles/r2d_*: noise generated between the loads and their first use; variants
mkdir -p gpurun_out
SPIRIT_B200_LIB=libSpirit.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "fused or iterate_block or single_steps or fullsize or 256" > gpurun_out/r2d_pytest.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2d_pytest.txt
tail -3 gpurun_out/r2d_pytest.txt
timeout 900 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit_fI.so" "LIB=libSpirit_fJ.so" "LIB=libSpirit_fK.so" "LIB=libSpirit_fL.so" "LIB=libSpirit.so" > gpurun_out/r2d_sweep.txt 2>&1
cat gpurun_out/r2d_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_fused -s 6 -c 1 -o gpurun_out/r2d_prof_main -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2d_ncu_main.log 2>&1
