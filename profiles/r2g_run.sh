#!/bin/bash
# GPU-box script of profiles/r2g_*: running-pointer march for interior CTAs (noise between loads and gradient), smoke, driver-like bench
mkdir -p gpurun_out
for lib in libSpirit.so libSpirit_fS.so; do
SPIRIT_B200_LIB=$lib timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "fused or iterate_block or single_steps or fullsize or 256" > gpurun_out/r2g_pytest_$lib.txt 2>&1; echo "pytest $lib exit $?" | tee -a gpurun_out/r2g_pytest_$lib.txt
tail -3 gpurun_out/r2g_pytest_$lib.txt
done
timeout 900 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit_fS.so" "LIB=libSpirit_fT.so" "LIB=libSpirit.so" "LIB=libSpirit_fS.so" > gpurun_out/r2g_sweep.txt 2>&1
cat gpurun_out/r2g_sweep.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2g_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 2200 gpurun_out/r2g_bench.json
SPIRIT_B200_LIB=libSpirit_fS.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6_fused -s 6 -c 1 -o gpurun_out/r2g_prof_fS -f python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2g_ncu_fS.log 2>&1
