#!/bin/bash
# GPU-box script of profiles/r2t_*: persistent a-passes with cp.async double buffering (k_ddi_fwd_a16p / k_ddi_inv_a16p) against the plain kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ddi_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short -x > gpurun_out/r2t_pytest_ddi.txt 2>&1; echo "pytest ddi exit $?" | tee -a gpurun_out/r2t_pytest_ddi.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2t_pytest_ddi.txt | head -30
for V in "SPIRIT_B200_FFT_PIPE_A=0 SPIRIT_B200_FFT_PIPE_B=0" "SPIRIT_B200_FFT_PIPE_A=1 SPIRIT_B200_FFT_PIPE_B=0" "SPIRIT_B200_FFT_PIPE_A=1 SPIRIT_B200_FFT_PIPE_B=1" "SPIRIT_B200_FFT_PIPE_A=1 SPIRIT_B200_FFT_PIPE_B=0 SPIRIT_B200_FFT_LG_A=2"  "SPIRIT_B200_FFT_PIPE_A=1 SPIRIT_B200_FFT_PIPE_B=0 SPIRIT_B200_FFT_LG_A=4"; do
  echo "== $V" | tee -a gpurun_out/r2t_sweep.txt
  env $V timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>&1 | grep -E "config" | cut -c90-200 | tee -a gpurun_out/r2t_sweep.txt
  env $V timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-120 | tee -a gpurun_out/r2t_sweep.txt
done
export SPIRIT_B200_FFT_PIPE_B=0
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r2t_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r2t_launches.log 2>&1
python profiles/launch_table.py gpurun_out/r2t_launches_c5_256.csv | tee gpurun_out/r2t_launch_table_c5_256_ddi.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 60 -c 5 --csv --log-file gpurun_out/r2t_launches_c3.csv python profiles/bench_configs.py c3 > gpurun_out/r2t_launches_c3.log 2>&1
python profiles/launch_table.py gpurun_out/r2t_launches_c3.csv | tee gpurun_out/r2t_launch_table_c3_ddi.txt
