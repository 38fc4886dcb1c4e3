#!/bin/bash
# GPU-box script of profiles/r1zz_*: closing run of round 1 -- whole GPU test suite, smoke, both bench arms, ncu --set full of the
# shipped DDI kernels at 256^3, launch list of the GNEB config
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1zz_pytest.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r1zz_pytest.txt
tail -3 gpurun_out/r1zz_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r1zz_smoke.txt
timeout 900 python bench.py > gpurun_out/r1zz_bench.json 2> gpurun_out/r1zz_bench.err; tail -c 1800 gpurun_out/r1zz_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r1zz_bench_reference.json 2> gpurun_out/r1zz_bench_reference.err; tail -c 600 gpurun_out/r1zz_bench_reference.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ddi_fwd_a|k_ddi_c_mult|k_ddi_inv_a|k_fft_pass" -s 30 -c 5 -o gpurun_out/r1zz_ddi256 -f python profiles/bench_c5.py --edge 256 --steps 2 > gpurun_out/r1zz_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_chain|k_reduce" -s 200 -c 40 --csv --log-file gpurun_out/r1zz_launches_c4_gneb.csv python profiles/bench_configs.py c4 > gpurun_out/r1zz_launches_c4.log 2>&1
timeout 300 python profiles/bench_configs.py c1 c3 c4 2>/dev/null | tee gpurun_out/r1zz_bench_configs.txt | cut -c1-160
timeout 200 python profiles/bench_c5.py --edge 256 --steps 10 2>/dev/null | grep config | tee gpurun_out/r1zz_bench_c5_256.txt | cut -c1-200
