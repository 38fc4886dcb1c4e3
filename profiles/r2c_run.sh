#!/bin/bash
# GPU-box script of profiles/r2c_*: fast interior march of the fused kernel (running pointers), per-warp progress counters
# instead of the CTA barrier, tile height 13 at 128 registers; parity tests for every variant that may ship
mkdir -p gpurun_out
for lib in libSpirit.so libSpirit_fF.so; do
  SPIRIT_B200_LIB=$lib timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "fused or iterate_block or single_steps or fullsize or 256" > gpurun_out/r2c_pytest_$lib.txt 2>&1; echo "pytest $lib exit $?" | tee -a gpurun_out/r2c_pytest_$lib.txt
  tail -3 gpurun_out/r2c_pytest_$lib.txt
done
timeout 900 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit_fF.so" "LIB=libSpirit_fG.so" "LIB=libSpirit_fH.so" "SPIRIT_B200_NO_FUSED=1" "LIB=libSpirit.so" "LIB=libSpirit_fF.so" > gpurun_out/r2c_sweep.txt 2>&1
cat gpurun_out/r2c_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_fused -s 6 -c 1 -o gpurun_out/r2c_prof_main -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c_ncu_main.log 2>&1
SPIRIT_B200_LIB=libSpirit_fF.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_fused -s 6 -c 1 -o gpurun_out/r2c_prof_fF -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c_ncu_fF.log 2>&1
