#!/bin/bash
# GPU-box script of profiles/r1u_*: fast DDI kernels instantiated per length; twiddles by look-up vs powers of one entry
mkdir -p gpurun_out
for L in libSpirit.so libSpirit_chain6.so; do
  echo "== $L" | tee -a gpurun_out/r1u_sweep.txt
  SPIRIT_B200_LIB=$L timeout 900 python -m pytest tests/test_ddi_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "ddi or dipolar" 2>&1 | tail -1 | tee -a gpurun_out/r1u_sweep.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-170 | tee -a gpurun_out/r1u_sweep.txt
  SPIRIT_B200_LIB=$L timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-120 | tee -a gpurun_out/r1u_sweep.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r1u_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1u_launches.log 2>&1
SPIRIT_B200_LIB=libSpirit_chain6.so timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r1u_launches_c5_256_chain6.csv python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1u_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 60 -c 5 --csv --log-file gpurun_out/r1u_launches_c3.csv python profiles/bench_configs.py c3 > gpurun_out/r1u_launches_c3.log 2>&1
