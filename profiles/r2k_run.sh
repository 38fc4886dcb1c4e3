#!/bin/bash
# GPU-box script of profiles/r2k_*: the driver's bench command with the sub-records of the other BASELINE configurations
mkdir -p gpurun_out
( time timeout 1200 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -4 gpurun_out/r2k_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2k_bench.json') if l.startswith('{')][-1])
print('value %.4e ms/step %.4f step frac %.3f kernel frac %.3f e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['step']['frac'], d['roofline']['frac'], d['e2e']['value']))
print(json.dumps(d['configs'], indent=1)[:6000])
print(d['cpu_baseline'])
PY
