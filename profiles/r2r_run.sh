#!/bin/bash
# GPU-box script of profiles/r2r_*: GNEB force options after the path-shortening race fix, RK4 over a chain against the restatement;
# dipolar tensor spectrum stored with its mirror symmetries (kc <= Pc/2, kb <= Pb/2): parity + per-launch tables
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gneb_gpu.py -m gpu -q --tb=short > gpurun_out/r2r_pytest_gneb.txt 2>&1; echo "pytest gneb exit $?" | tee -a gpurun_out/r2r_pytest_gneb.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2r_pytest_gneb.txt | head -30
timeout 1200 python -m pytest tests/test_ddi_gpu.py tests/test_fullsize_gpu.py tests/test_reference_inputs_gpu.py -m gpu -q --tb=short > gpurun_out/r2r_pytest_ddi.txt 2>&1; echo "pytest ddi exit $?" | tee -a gpurun_out/r2r_pytest_ddi.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2r_pytest_ddi.txt | head -30
for M in 1 0; do
  echo "== SPIRIT_B200_DDI_MIRROR=$M" | tee -a gpurun_out/r2r_sweep.txt
  SPIRIT_B200_DDI_MIRROR=$M SPIRIT_B200_DDI_VERBOSE=1 timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>&1 | grep -E "config|mirror" | cut -c1-400 | tee -a gpurun_out/r2r_sweep.txt
  SPIRIT_B200_DDI_MIRROR=$M timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-300 | tee -a gpurun_out/r2r_sweep.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r2r_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r2r_launches.log 2>&1
python profiles/launch_table.py gpurun_out/r2r_launches_c5_256.csv | tee gpurun_out/r2r_launch_table_c5_256_ddi.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 60 -c 5 --csv --log-file gpurun_out/r2r_launches_c3.csv python profiles/bench_configs.py c3 > gpurun_out/r2r_launches_c3.log 2>&1
python profiles/launch_table.py gpurun_out/r2r_launches_c3.csv | tee gpurun_out/r2r_launch_table_c3_ddi.txt
