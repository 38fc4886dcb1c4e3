#!/bin/bash
# GPU-box script of profiles/r2o_*: new tests (thermal variates, verbatim reference inputs, reduced configs[2], anisotropy tables),
# T > 0 statistics with the 32-bit radius uniforms, bench line with the fp64-variates check
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_thermal_variates_gpu.py tests/test_reference_inputs_gpu.py tests/test_ddi_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "variates or moments or distribution or tails or verbatim or gaussian or reduced_size or anisotropy_table or thermal or langevin or fused" > gpurun_out/r2o_pytest.txt 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r2o_pytest.txt
tail -6 gpurun_out/r2o_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2o_bench.json') if l.startswith('{')][-1])
print('ms/step %.4f step frac %.3f' % (d['ms_per_step'], d['roofline']['step']['frac']), d['checks'])
PY
