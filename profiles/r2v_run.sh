#!/bin/bash
# GPU-box script of profiles/r2v_*: second granularity sweep of the dipolar passes (plain kernels), 256^3 and 512^3; STT tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short -k "spin_transfer" > gpurun_out/r2v_pytest_stt.txt 2>&1; echo "pytest stt exit $?" | tee -a gpurun_out/r2v_pytest_stt.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2v_pytest_stt.txt | head -20
run() { E=$1; shift; echo "== $E $*" | tee -a gpurun_out/r2v_sweep.txt; env SPIRIT_B200_FFT_PIPE_A=0 SPIRIT_B200_FFT_PIPE_B=0 "$@" timeout 300 python profiles/bench_c5.py --edge $E --steps 10 2>&1 | grep -E "config" | cut -c90-150 | tee -a gpurun_out/r2v_sweep.txt; }
P=SPIRIT_B200_FFT
run 256 ${P}_LG_A=1 ${P}_LG_B=1 ${P}_SEQ_B=0
run 256 ${P}_LG_A=0 ${P}_LG_B=1 ${P}_SEQ_B=0
run 256 ${P}_LG_A=1 ${P}_LG_B=0 ${P}_SEQ_B=1
run 256 ${P}_LG_A=1 ${P}_LG_B=0 ${P}_SEQ_B=0
run 256 ${P}_LG_A=1 ${P}_LG_B=2 ${P}_SEQ_B=0
run 256 ${P}_LG_A=1 ${P}_LG_B=1 ${P}_SEQ_B=0 ${P}_LG_C=1
run 512 ${P}_LG_A=2
run 512 ${P}_LG_A=1
run 512 ${P}_LG_A=0
run 512 ${P}_LG_A=1 ${P}_LG_B=0 ${P}_SEQ_B=1
run 512 ${P}_LG_A=1 ${P}_LG_B=0 ${P}_SEQ_B=0
run 512 ${P}_LG_A=1 ${P}_LG_B=1 ${P}_SEQ_B=0
run 512 ${P}_LG_A=1 ${P}_LG_C=0
