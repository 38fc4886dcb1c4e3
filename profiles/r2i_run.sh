#!/bin/bash
# GPU-box script of profiles/r2i_* (2 GPUs): slabs with the fused kernel (two halo planes, one exchange per iteration): parity
# against one GPU, weak-scaling bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -k "world0 or 2" > gpurun_out/r2i_pytest.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2i_pytest.txt
grep -E "OK|FAIL|MGPU|passed|failed" gpurun_out/r2i_pytest.txt | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; tail -c 1500 gpurun_out/r2i_bench_n2.json
