#!/bin/bash
# GPU-box script of profiles/r2q_*: GNEB force options (energy-weighted springs, path shortening, moving / translating endpoints),
# RK4 over a chain, GNEB with the dipolar convolution; lazy effective-field mirror; then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gneb_gpu.py -m gpu -q --tb=short > gpurun_out/r2q_pytest_gneb.txt 2>&1; echo "pytest gneb exit $?" | tee -a gpurun_out/r2q_pytest_gneb.txt
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/r2q_pytest_gneb.txt | head -60
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gneb_gpu.py > gpurun_out/r2q_pytest_all.txt 2>&1; echo "pytest all exit $?" | tee -a gpurun_out/r2q_pytest_all.txt
tail -8 gpurun_out/r2q_pytest_all.txt
