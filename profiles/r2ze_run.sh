#!/bin/bash
# GPU-box script of profiles/r2ze_*: dipolar passes with register-resident first / last stages (k_ddi_c_mult16f: forward last stage ->
# tensor multiply -> inverse first stage in registers; k_fft_pass16r) against the round's kernels: parity, then timings per switch
mkdir -p gpurun_out
O=gpurun_out/r2ze_sweep.txt; : > $O
timeout 600 python -m pytest tests/test_ddi_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2ze_pytest_ddi.txt
run() { # label, env...
  echo "== $1" | tee -a $O; shift
  env "$@" timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-150 | tee -a $O
  env "$@" timeout 300 python profiles/bench_c5.py --edge 512 --steps 5 2>/dev/null | grep config | cut -c90-150 | tee -a $O
  env "$@" timeout 300 python profiles/bench_configs.py c3 2>/dev/null | head -1 | cut -c1-120 | tee -a $O
}
run "round-2 kernels (C_FUSED=0 PASS_REG=0)" SPIRIT_B200_DDI_C_FUSED=0 SPIRIT_B200_FFT_PASS_REG=0
run "c fused only" SPIRIT_B200_FFT_PASS_REG=0
run "b register stages only" SPIRIT_B200_DDI_C_FUSED=0
run "both (default)" X=1
run "both, c with one column less per CTA (LG_C one below default: 256^3 lg 0, 512^3 n/a)" SPIRIT_B200_FFT_LG_C=0
run "both, libSpirit_cf256 (c: 256 threads, 2 CTAs / SM, 128 registers)" SPIRIT_B200_LIB=libSpirit_cf256.so
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r2ze_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > /dev/null 2>&1
python profiles/launch_table.py gpurun_out/r2ze_launches_c5_256.csv | tee gpurun_out/r2ze_launch_table_c5_256_ddi.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r2ze_launches_c5_512.csv python profiles/bench_c5.py --edge 512 --steps 2 > /dev/null 2>&1
python profiles/launch_table.py gpurun_out/r2ze_launches_c5_512.csv | tee gpurun_out/r2ze_launch_table_c5_512_ddi.txt
