#!/bin/bash
# GPU-box script of profiles/r2za_* (2 GPUs): ka pencils with the transposes on the copy engines (2-D peer copies per component under
# the passes of the other components): parity, then 512^3 + DDI SIB against the other schedules
mkdir -p gpurun_out
SPIRIT_B200_DDI_PENCIL_DMA=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tests/mgpu_worker.py > gpurun_out/r2za_mgpu_n2.txt 2>&1; echo "worker exit $?" | tee -a gpurun_out/r2za_mgpu_n2.txt
grep -E "DDI|FAIL|MGPU|Error|error" gpurun_out/r2za_mgpu_n2.txt | cut -c1-200 | tail -14
run() { echo "== $*" | tee -a gpurun_out/r2za_sweep.txt; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 profiles/bench_c5.py --edge 512 --steps 5 2>gpurun_out/r2za_err.txt | grep config | cut -c90-160 | tee -a gpurun_out/r2za_sweep.txt; grep -i "error\|Traceback" gpurun_out/r2za_err.txt | head -3; }
run SPIRIT_B200_DDI_PENCIL_DMA=1
run SPIRIT_B200_DDI_PENCIL_CTAS=4
run SPIRIT_B200_DDI_PENCIL_CTAS=5
run SPIRIT_B200_DDI_PENCIL_CTAS=3
