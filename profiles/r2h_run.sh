#!/bin/bash
# GPU-box script of profiles/r2h_*: 32-bit plane arithmetic on the uniform datapath, hook gradient through shared memory
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q -k "fused or iterate_block or single_steps or fullsize or 256" > gpurun_out/r2h_pytest.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2h_pytest.txt
tail -3 gpurun_out/r2h_pytest.txt
timeout 900 python profiles/sweep.py "LIB=libSpirit.so" "LIB=libSpirit.so" > gpurun_out/r2h_sweep.txt 2>&1
cat gpurun_out/r2h_sweep.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2h_bench.json') if l.startswith('{')][-1]); print('bench 20/5: ms/step', d['ms_per_step'], 'step frac', d['roofline']['step']['frac'], 'kernel frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
