#!/bin/bash
# GPU-box script of profiles/r2zi_*: twiddles of the radix-8 stages from look-ups (first stage: w, w^2, w^4 + four products; later
# stages: the table) against powers multiplied up from one entry (libSpirit_tw0.so, -DSB_FFT_TW_MODE=0)
mkdir -p gpurun_out
O=gpurun_out/r2zi_sweep.txt; : > $O
timeout 600 python -m pytest tests/test_ddi_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2zi_pytest_ddi.txt
run() { # label, env...
  echo "== $1" | tee -a $O; shift
  env "$@" timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-150 | tee -a $O
  env "$@" timeout 300 python profiles/bench_c5.py --edge 512 --steps 5 2>/dev/null | grep config | cut -c90-150 | tee -a $O
  env "$@" timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-120 | tee -a $O
}
run "look-ups (product)" X=1
run "powers multiplied up (libSpirit_tw0.so)" SPIRIT_B200_LIB=libSpirit_tw0.so
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r2zi_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > /dev/null 2>&1
python profiles/launch_table.py gpurun_out/r2zi_launches_c5_256.csv | tee gpurun_out/r2zi_launch_table_c5_256_ddi.txt
