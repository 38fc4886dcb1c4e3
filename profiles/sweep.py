#!/usr/bin/env python
"""Tuning sweep (GPU box): bench.py device-resident leg under different block shapes / march lengths / build variants.
usage: python profiles/sweep.py "LIB=libSpirit.so BX=128 BY=4 LC=16" "LIB=... BX=.." ...   (any subset of keys)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=")
        if k == "LIB":
            env["SPIRIT_B200_LIB"] = v
        elif k == "ARGS":
            pass
        elif k.startswith("SPIRIT_"):
            env[k] = v
        elif k == "FUSED_LC":
            env["SPIRIT_B200_FUSED_LC"] = v
        else:
            env["SPIRIT_B200_SC6_" + k] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "100", "--warmup", "10", "--no-e2e", "--no-cpu-baseline", "--no-extras"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        rf = d["roofline"]
        kernels = rf["stage_ms"] if "stage_ms" in rf else [rf["kernel_ms"]]  # two-pass stages, or the one fused kernel
        print("%-50s %.4e spin-steps/s  %.4f ms/step  kernel(s) %s ms  frac %.3f" % (
            spec, d["value"], d["ms_per_step"], ["%.4f" % x for x in kernels], rf["step"]["frac"]), flush=True)
    except Exception as e:  # noqa: BLE001
        print(spec, "FAILED", e, r.stderr[-500:], flush=True)
