#!/bin/bash
# GPU-box script of profiles/r1p_*: fast DDI passes with all loads of a thread in flight; columns-per-CTA sweep at 256^3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddi_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "ddi or dipolar" > gpurun_out/r1p_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1p_pytest.txt
grep -v "^    \|^  \|^$\|^2026\|^====" gpurun_out/r1p_pytest.txt | tail -6
run() { echo "== $*" | tee -a gpurun_out/r1p_sweep.txt; env "$@" timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | cut -c90-200 | tee -a gpurun_out/r1p_sweep.txt; }
run X=0
run SPIRIT_B200_FFT_LG_C=1
run SPIRIT_B200_FFT_LG_C=0
run SPIRIT_B200_FFT_LG_B=1
run SPIRIT_B200_FFT_LG_B=3
run SPIRIT_B200_FFT_LG_A=2
run SPIRIT_B200_FFT_LG_A=4
timeout 600 python profiles/bench_configs.py c3 2>/dev/null | head -2 | tee gpurun_out/r1p_bench_c3.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r1p_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1p_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 -o gpurun_out/r1p_ddi256 -f python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1p_ncu.log 2>&1
