#!/bin/bash
# GPU-box script of profiles/r1s_*: radix-4 in-register stages (FFT_E = 4; 64 registers, two 512-thread CTAs per SM) against radix-8
mkdir -p gpurun_out
SPIRIT_B200_LIB=libSpirit_e4m2.so timeout 900 python -m pytest tests/test_ddi_gpu.py -m gpu -x -q > gpurun_out/r1s_pytest_e4m2.txt 2>&1; tail -2 gpurun_out/r1s_pytest_e4m2.txt
for L in libSpirit.so libSpirit_e4m2.so libSpirit_e4m1.so; do
  echo "== $L" | tee -a gpurun_out/r1s_sweep.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-170 | tee -a gpurun_out/r1s_sweep.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-120 | tee -a gpurun_out/r1s_sweep.txt
done
SPIRIT_B200_LIB=libSpirit_e4m2.so timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r1s_launches_c5_256_e4m2.csv python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1s_launches.log 2>&1
