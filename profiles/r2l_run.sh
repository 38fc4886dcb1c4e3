#!/bin/bash
# GPU-box script of profiles/r2l_*: whole GPU test suite after the fused / peer / bench changes, smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2l_pytest.txt
tail -5 gpurun_out/r2l_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2l_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('ms/step', d['ms_per_step'], 'c1', d['configs']['c1'])"
