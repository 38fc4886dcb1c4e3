#!/bin/bash
# GPU-box script of profiles/r1n_*: where the dipolar convolution spends its time at 256^3 (C5 on one GPU) and at 2048x2048x4 (C3)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass|k_sc6|k_llg" -s 30 -c 40 --csv --log-file gpurun_out/r1n_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1n_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 -o gpurun_out/r1n_ddi256 -f python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1n_ncu.log 2>&1
tail -2 gpurun_out/r1n_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 60 -c 5 -o gpurun_out/r1n_ddi_c3 -f python profiles/bench_configs.py c3 > gpurun_out/r1n_ncu_c3.log 2>&1
tail -2 gpurun_out/r1n_ncu_c3.log
timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | tee gpurun_out/r1n_bench_c5_256.txt
