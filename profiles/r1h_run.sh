#!/bin/bash
# GPU-box script of profiles/r1h_*: whole GPU test suite, bench, launch list, ncu capture of the stage kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1h_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1h_pytest.txt
tail -4 gpurun_out/r1h_pytest.txt
timeout 900 python profiles/sweep.py "LIB=libSpirit_base.so" "LIB=libSpirit.so" "LIB=libSpirit_base.so" "LIB=libSpirit.so" > gpurun_out/r1h_sweep.txt 2>&1
cat gpurun_out/r1h_sweep.txt
timeout 900 python bench.py > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err; tail -c 3000 gpurun_out/r1h_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sc6 -s 60 -c 2 -o gpurun_out/r1h_prof -f python bench.py --steps 5 --warmup 30 --no-e2e --no-cpu-baseline > gpurun_out/r1h_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1h_launches.log 2>&1
tail -2 gpurun_out/r1h_ncu.log
