#!/bin/bash
# GPU-box script of profiles/r2zh_*: pinned sites and defects against the reference built with its two compile-time options,
# then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pinning_defects_gpu.py -m gpu -q --tb=short > gpurun_out/r2zh_pytest_pd.txt 2>&1; echo "exit $?" >> gpurun_out/r2zh_pytest_pd.txt
tail -40 gpurun_out/r2zh_pytest_pd.txt
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_pinning_defects_gpu.py > gpurun_out/r2zh_pytest_all.txt 2>&1; echo "exit $?" >> gpurun_out/r2zh_pytest_all.txt
tail -8 gpurun_out/r2zh_pytest_all.txt
