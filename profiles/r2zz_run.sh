#!/bin/bash
# GPU-box script of profiles/r2zz_*: closing run of round 2 -- whole GPU test suite, smoke, both bench arms, launch list of the bench
# command, ncu --set full of the fused kernel, the other BASELINE configurations, launch tables of the dipolar passes
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2zz_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2zz_pytest.txt
tail -3 gpurun_out/r2zz_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee gpurun_out/r2zz_smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; tail -c 600 gpurun_out/r2zz_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2zz_bench_reference.json 2> gpurun_out/r2zz_bench_reference.err; tail -c 700 gpurun_out/r2zz_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2zz_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2zz_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sc6_fused" -s 20 -c 2 -o gpurun_out/r2zz_fused -f python bench.py --steps 30 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > gpurun_out/r2zz_ncu_fused.log 2>&1
timeout 300 python profiles/bench_configs.py c1 c4 2>/dev/null | cut -c1-400 | tee gpurun_out/r2zz_configs_c1_c4.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 --csv --log-file gpurun_out/r2zz_launches_c5_256.csv python profiles/bench_c5.py --edge 256 --steps 2 > /dev/null 2>&1
python profiles/launch_table.py gpurun_out/r2zz_launches_c5_256.csv | tee gpurun_out/r2zz_launch_table_c5_256_ddi.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_ddi|k_fft_pass" -s 60 -c 5 --csv --log-file gpurun_out/r2zz_launches_c3.csv python profiles/bench_configs.py c3 > /dev/null 2>&1
python profiles/launch_table.py gpurun_out/r2zz_launches_c3.csv | tee gpurun_out/r2zz_launch_table_c3_ddi.txt
