#!/bin/bash
# GPU-box script of profiles/r1y_*: b-pass with two column groups per CTA (full sectors for the long transforms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddi_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "ddi or dipolar" 2>&1 | tail -1 | tee gpurun_out/r1y_pytest.txt
SPIRIT_B200_FFT_SEQ_B=1 timeout 900 python -m pytest tests/test_ddi_gpu.py -m gpu -x -q 2>&1 | tail -1 | tee -a gpurun_out/r1y_pytest.txt
run() { echo "== $*" | tee -a gpurun_out/r1y_sweep.txt; env "$@" timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-170 | tee -a gpurun_out/r1y_sweep.txt; env "$@" timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-110 | tee -a gpurun_out/r1y_sweep.txt; }
run SPIRIT_B200_FFT_SEQ_B=0
run SPIRIT_B200_FFT_SEQ_B=1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 60 -c 5 --csv --log-file gpurun_out/r1y_launches_c3.csv python profiles/bench_configs.py c3 > gpurun_out/r1y_launches_c3.log 2>&1
