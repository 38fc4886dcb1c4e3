#!/bin/bash
# GPU-box script of profiles/r2a_*: first run of the fused predictor + corrector kernel (sc6_fused.cuh): its parity tests,
# A/B of the build variants against the two-pass kernels, launch list + ncu --set full of the default variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fused or iterate_block or single_steps" > gpurun_out/r2a_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.txt
tail -15 gpurun_out/r2a_pytest.txt
timeout 900 python profiles/sweep.py "SPIRIT_B200_NO_FUSED=1" "LIB=libSpirit.so" "LIB=libSpirit_fB.so" "LIB=libSpirit_fC.so" "LIB=libSpirit_fD.so" "LIB=libSpirit.so FUSED_LC=16" "LIB=libSpirit.so FUSED_LC=64" "LIB=libSpirit_fB.so FUSED_LC=16" "LIB=libSpirit.so" > gpurun_out/r2a_sweep.txt 2>&1
cat gpurun_out/r2a_sweep.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 2500 gpurun_out/r2a_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_fused -s 20 -c 1 -o gpurun_out/r2a_prof -f python bench.py --steps 5 --warmup 10 --no-e2e --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
tail -2 gpurun_out/r2a_ncu.log
