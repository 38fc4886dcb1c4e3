#!/bin/bash
# GPU-box script of profiles/r1t_*: whole GPU test suite, smoke, bench (both arms), launch lists of bench.py and of the DDI configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1t_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1t_pytest.txt
tail -4 gpurun_out/r1t_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r1t_smoke.txt
timeout 900 python bench.py > gpurun_out/r1t_bench.json 2> gpurun_out/r1t_bench.err; tail -c 2500 gpurun_out/r1t_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r1t_bench_reference.json 2> gpurun_out/r1t_bench_reference.err; tail -c 1200 gpurun_out/r1t_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1t_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r1t_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 60 -c 10 --csv --log-file gpurun_out/r1t_launches_c3.csv python profiles/bench_configs.py c3 > gpurun_out/r1t_launches_c3.log 2>&1
timeout 600 python profiles/bench_configs.py c1 c3 c4 2>/dev/null | tee gpurun_out/r1t_bench_configs.txt | cut -c1-200
