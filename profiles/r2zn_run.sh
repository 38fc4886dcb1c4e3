#!/bin/bash
# GPU-box script of profiles/r2zn_* (2 GPUs): configs[3] alone on a sharded chain, this library against the one before the pinning /
# defects change (the N = 2 bench line of r2zm showed 721 iterations/s where the second session had 3 813)
mkdir -p gpurun_out
for L in libSpirit.so libSpirit_prepin.so; do
  echo "== $L" | tee -a gpurun_out/r2zn_c4_n2.txt
  SPIRIT_B200_LIB=$L timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 profiles/c4_only.py 2>/dev/null | grep iterations_per_s | tee -a gpurun_out/r2zn_c4_n2.txt
done
