#!/bin/bash
# GPU-box script of profiles/r1x_*: marching kernels specialised for the plain Hamiltonian (SC6_PLAIN) against the unspecialised ones
mkdir -p gpurun_out
timeout 900 python profiles/sweep.py "SPIRIT_B200_SC6_PLAIN=0" "SPIRIT_B200_SC6_PLAIN=1" "SPIRIT_B200_SC6_PLAIN=0" "SPIRIT_B200_SC6_PLAIN=1" > gpurun_out/r1x_sweep.txt 2>&1
cat gpurun_out/r1x_sweep.txt
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r1x_pytest.txt
