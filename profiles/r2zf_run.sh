#!/bin/bash
# GPU-box script of profiles/r2zf_*: b-pass with register-resident stages and the load-phase thread mapping in the last stage
# (two column groups per CTA), c-pass shapes per length; ncu --set full of the register-resident c-pass
mkdir -p gpurun_out
O=gpurun_out/r2zf_sweep.txt; : > $O
timeout 600 python -m pytest tests/test_ddi_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2zf_pytest_ddi.txt
run() { # label, env...
  echo "== $1" | tee -a $O; shift
  env "$@" timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-150 | tee -a $O
  env "$@" timeout 300 python profiles/bench_c5.py --edge 512 --steps 5 2>/dev/null | grep config | cut -c90-150 | tee -a $O
  env "$@" timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-120 | tee -a $O
}
run "default (c register-resident, b register stages at every length)" X=1
run "b-pass: round-2 kernel (PASS_REG=0)" SPIRIT_B200_FFT_PASS_REG=0
run "b-pass: one column group per CTA (SEQ_B=0)" SPIRIT_B200_FFT_SEQ_B=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ddi_c_mult16f|k_fft_pass16r" -s 6 -c 3 -o gpurun_out/r2zf_ddi256 -f python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r2zf_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r2zf_launches_c5_512.csv python profiles/bench_c5.py --edge 512 --steps 2 > /dev/null 2>&1
python profiles/launch_table.py gpurun_out/r2zf_launches_c5_512.csv | tee gpurun_out/r2zf_launch_table_c5_512_ddi.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 60 -c 5 --csv --log-file gpurun_out/r2zf_launches_c3.csv python profiles/bench_configs.py c3 > /dev/null 2>&1
python profiles/launch_table.py gpurun_out/r2zf_launches_c3.csv | tee gpurun_out/r2zf_launch_table_c3_ddi.txt
