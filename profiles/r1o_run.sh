#!/bin/bash
# GPU-box script of profiles/r1o_*: fast DDI passes (in-place power-of-two FFT kernels, half-length real a-pass, tiled real tensor)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddi_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "ddi or dipolar" > gpurun_out/r1o_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1o_pytest.txt
grep -v "^    \|^  \|^$\|^2026\|^====" gpurun_out/r1o_pytest.txt | tail -12
timeout 600 python profiles/bench_configs.py c3 2>/dev/null | head -2 | tee gpurun_out/r1o_bench_c3.txt
timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | tee gpurun_out/r1o_bench_c5_256.txt
SPIRIT_B200_FFT_FAST=0 timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | tee gpurun_out/r1o_bench_c5_256_old.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 10 --csv --log-file gpurun_out/r1o_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1o_launches.log 2>&1
