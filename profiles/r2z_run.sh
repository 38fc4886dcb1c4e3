#!/bin/bash
# GPU-box script of profiles/r2z_* (2 GPUs): ka pencils with the per-component overlapped schedule (persistent a-passes on a bounded
# number of CTAs under the b- / c-passes of the other components): parity, then 512^3 + DDI SIB for several CTA bounds
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tests/mgpu_worker.py > gpurun_out/r2z_mgpu_n2.txt 2>&1; echo "worker exit $?" | tee -a gpurun_out/r2z_mgpu_n2.txt
grep -E "DDI|FAIL|MGPU|Error|error" gpurun_out/r2z_mgpu_n2.txt | cut -c1-200 | tail -14
run() { echo "== $*" | tee -a gpurun_out/r2z_sweep.txt; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 profiles/bench_c5.py --edge 512 --steps 5 2>gpurun_out/r2z_err.txt | grep config | cut -c90-160 | tee -a gpurun_out/r2z_sweep.txt; grep -i "error\|Traceback" gpurun_out/r2z_err.txt | head -3; }
run SPIRIT_B200_DDI_PENCIL_OVERLAP=0
run SPIRIT_B200_DDI_PENCIL_CTAS=2
run SPIRIT_B200_DDI_PENCIL_CTAS=4
run SPIRIT_B200_DDI_PENCIL_CTAS=8
run SPIRIT_B200_DDI_PENCIL_CTAS=16
