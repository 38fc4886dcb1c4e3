#!/bin/bash
# GPU-box script of profiles/r2zb_* (8 GPUs): 512^3 + DDI SIB with ka pencils: transposes on the copy engines vs the overlapped
# push / pull schedule; then the bench line at N = 8 with the faster of the two
mkdir -p gpurun_out
run() { echo "== $*" | tee -a gpurun_out/r2zb_sweep.txt; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 profiles/bench_c5.py --edge 512 --steps 8 2>gpurun_out/r2zb_err.txt | grep config | cut -c90-160 | tee -a gpurun_out/r2zb_sweep.txt; grep -i "error\|Traceback" gpurun_out/r2zb_err.txt | head -3; }
run SPIRIT_B200_DDI_PENCIL_DMA=1
run SPIRIT_B200_DDI_PENCIL_DMA=0
run SPIRIT_B200_DDI_PENCIL_DMA=1 SPIRIT_B200_FFT_LG_A=1
BEST=$(python - <<'PY'
import re
t=open('gpurun_out/r2zb_sweep.txt').read().split('== ')[1:]
ms=[float(re.search(r'ms_per_iteration": ([0-9.]+)', x).group(1)) for x in t[:2]]
print(1 if ms[0] < ms[1] else 0)
PY
)
echo "bench with SPIRIT_B200_DDI_PENCIL_DMA=$BEST" | tee -a gpurun_out/r2zb_sweep.txt
SPIRIT_B200_DDI_PENCIL_DMA=$BEST timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2zb_bench_n8.json 2> gpurun_out/r2zb_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2zb_bench_n8.json') if l.startswith('{')][-1])
print('ms/step %.4f value %.4g' % (d['ms_per_step'], d['value']), 'e2e', d['e2e']['value'], d['clocks'])
print('c4', {k: d['configs']['c4'].get(k) for k in ('iterations_per_s', 'barrier_meV', 'error')})
print('c5', {k: d['configs']['c5'].get(k) for k in ('ms_per_iteration', 'error')})
print('parity', d['multi_gpu_parity'])
PY
tail -3 gpurun_out/r2zb_bench_n8.err | cut -c1-300
