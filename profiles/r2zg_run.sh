#!/bin/bash
# GPU-box script of profiles/r2zg_*: c-pass with the next component's loads in flight and the CTA's tensor block staged in shared
# memory by cp.async (six components when the occupancy calculator still gives three CTAs per SM)
mkdir -p gpurun_out
O=gpurun_out/r2zg_sweep.txt; : > $O
timeout 600 python -m pytest tests/test_ddi_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2zg_pytest_ddi.txt
run() { # label, env...
  echo "== $1" | tee -a $O; shift
  env "$@" timeout 300 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | cut -c90-150 | tee -a $O
  env "$@" timeout 300 python profiles/bench_c5.py --edge 512 --steps 5 2>/dev/null | grep config | cut -c90-150 | tee -a $O
}
run "default (tensor block staged where it fits)" X=1
run "tensor through __ldg (STAGE_TENSOR=0)" SPIRIT_B200_DDI_C_STAGE_TENSOR=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ddi_c_mult16f" -s 2 -c 1 -o gpurun_out/r2zg_ddi256 -f python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r2zg_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r2zg_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > /dev/null 2>&1
python profiles/launch_table.py gpurun_out/r2zg_launches_c5_256.csv | tee gpurun_out/r2zg_launch_table_c5_256_ddi.txt
