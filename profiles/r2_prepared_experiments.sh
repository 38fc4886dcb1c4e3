#!/bin/bash
# Prepared for round 2 (not run in round 1: the GPU budget was spent). Build the variants HERE first (CPU, ~2 min each):
#   SPIRIT_B200_VARIANT=e16       SPIRIT_B200_DEFINES="-DSB_FFT_LG_E=4"                                  python -m spirit_b200.build
#   SPIRIT_B200_VARIANT=plain_a   SPIRIT_B200_DEFINES="-DSB_SC6_NO_STT=1 -DSB_SC6_NO_ANISO_FULL=1"      python -m spirit_b200.build
#   SPIRIT_B200_VARIANT=plain_b   SPIRIT_B200_DEFINES="-DSB_SC6_NO_STT=1 -DSB_SC6_NO_EXTRAS=1"          python -m spirit_b200.build
# then: gpurun --timeout 1500 -- 'bash profiles/r2_prepared_experiments.sh'
mkdir -p gpurun_out
# 1. marching kernels with two of the three uniform tests compiled out (profiles/r1zz_ptxas_study_plain_variants.txt)
timeout 900 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit_plain_a.so" "LIB=libSpirit_plain_b.so" "LIB=libSpirit.so" "LIB=libSpirit_plain_a.so" "LIB=libSpirit_plain_b.so" \
    > gpurun_out/r2a_sweep_plain_variants.txt 2>&1
cat gpurun_out/r2a_sweep_plain_variants.txt
# 2. radix-16 FFT stages: parity first, then the two DDI configurations
SPIRIT_B200_LIB=libSpirit_e16.so timeout 900 python -m pytest tests/test_ddi_gpu.py -m gpu -x -q 2>&1 | tail -1 | tee gpurun_out/r2a_pytest_e16.txt
for L in libSpirit.so libSpirit_e16.so; do
  echo "== $L" | tee -a gpurun_out/r2a_sweep_e16.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-170 | tee -a gpurun_out/r2a_sweep_e16.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-120 | tee -a gpurun_out/r2a_sweep_e16.txt
done
