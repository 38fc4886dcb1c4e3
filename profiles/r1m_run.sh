#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddi_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "ddi or dipolar" > gpurun_out/r1m_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1m_pytest.txt
grep -v "^    \|^  \|^$\|^2026\|^====" gpurun_out/r1m_pytest.txt | tail -12
timeout 600 python profiles/bench_configs.py c3 2>/dev/null | tee gpurun_out/r1m_bench_c3.txt
SPIRIT_B200_FFT16=0 timeout 600 python profiles/bench_configs.py c3 2>/dev/null | head -2 | tee gpurun_out/r1m_bench_c3_old.txt
