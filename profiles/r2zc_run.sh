#!/bin/bash
# GPU-box script of profiles/r2zc_*: chain gradient by index arithmetic on nearest-neighbour lattices (GNEB tests + C4), radix-16
# build variant of the dipolar passes (libSpirit_e4.so: -DSB_FFT_LG_E=4 -DSB_FFT_THREADS=256) against the product
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gneb_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short > gpurun_out/r2zc_pytest_gneb.txt 2>&1; echo "pytest gneb exit $?" | tee -a gpurun_out/r2zc_pytest_gneb.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2zc_pytest_gneb.txt | head -20
timeout 300 python profiles/bench_configs.py c4 2>/dev/null | head -2 | cut -c1-300 | tee gpurun_out/r2zc_c4.txt
for L in libSpirit.so libSpirit_e4.so; do
  echo "== $L" | tee -a gpurun_out/r2zc_sweep.txt
  SPIRIT_B200_LIB=$L timeout 600 python -m pytest tests/test_ddi_gpu.py -m gpu -x -q 2>&1 | tail -1 | tee -a gpurun_out/r2zc_sweep.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-150 | tee -a gpurun_out/r2zc_sweep.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_c5.py --edge 512 --steps 5 2>/dev/null | grep config | cut -c90-150 | tee -a gpurun_out/r2zc_sweep.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-120 | tee -a gpurun_out/r2zc_sweep.txt
done
