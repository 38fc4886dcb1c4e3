#!/usr/bin/env python
"""Secondary measurements on the other BASELINE.json configurations (not bench.py's headline line):
  C1 100x100x1 default input.cfg, Depondt             (launch-latency regime)
  C3 2048x2048x4 thin film + DDI (FFT), VP and Depondt (dipolar convolution: 1008 B per spin and gradient evaluation model)
  C4 GNEB 64 images of 256x256x1, VP                   (image-batched kernels)
usage: python profiles/bench_configs.py [c1] [c3] [c4]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bench import fill_random  # noqa: E402
from spirit_b200 import capi, session as S  # noqa: E402
from tests import cfgs  # noqa: E402

lib = capi.load_product()
tmp = tempfile.mkdtemp()
which = sys.argv[1:] or ["c1", "c3", "c4"]


def cfg(name, preset, **kw):
    path = os.path.join(tmp, name + ".cfg")
    open(path, "w").write(cfgs.render(preset, **kw))
    return path


if "c1" in which:
    p = S.Session(lib, cfg("c1", "default"))
    p.plus_z()
    p.skyrmion(5.0, phase=-90.0)
    p.upload()
    p.iterate_device(S.SOLVER_DEPONDT, 200)
    n = 5000
    ms = p.iterate_device(S.SOLVER_DEPONDT, n)
    print(json.dumps({"config": "C1 100x100x1 Depondt", "iterations_per_s": n / ms * 1e3, "spin_steps_per_s": p.nos * n / ms * 1e3,
                      "us_per_iteration": ms / n * 1e3}), flush=True)
    p.close()

if "c3" in which:
    t0 = time.time()
    p = S.Session(lib, cfg("c3", "cubic256", n_basis_cells="2048 2048 4", boundary_conditions="0 0 0", ddi_method="fft",
                           ddi_n_periodic_images="0 0 0", external_field_magnitude=25, anisotropy_magnitude=0, llg_temperature=0))
    fill_random(p)
    t1 = time.time()
    p.upload()  # builds the plan: tensor + its spectrum
    t2 = time.time()
    for solver, name, bytes_per in ((S.SOLVER_VP, "VP", 144 + 1008), (S.SOLVER_DEPONDT, "Depondt", 120 + 2 * 1008)):
        p.iterate_device(solver, 3)
        n = 20
        ms = p.iterate_device(solver, n)
        rate = p.nos * n / ms * 1e3
        print(json.dumps({"config": "C3 2048x2048x4 DDI-FFT " + name, "ms_per_iteration": ms / n, "spin_steps_per_s": rate,
                          "model_bytes_per_spin_step": bytes_per, "model_GBps": rate * bytes_per / 1e9,
                          "setup_s": {"state+random": t1 - t0, "upload+ddi_plan": t2 - t1}}), flush=True)
    p.close()
    # cuFFT (through torch.fft) timed alongside as a CHECK only: the library-style, un-pruned execution of the same
    # convolution (3 forward R2C + 3 inverse C2R of the padded 8 x 4096 x 4096 lattice, multiply not included)
    try:
        import torch
        x = torch.zeros((3, 8, 4096, 4096), dtype=torch.float64, device="cuda")
        x[:, :4, :2048, :2048] = 1.0
        for _ in range(2):
            y = torch.fft.rfftn(x, dim=(1, 2, 3))
            z = torch.fft.irfftn(y, s=(8, 4096, 4096), dim=(1, 2, 3))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = torch.fft.rfftn(x, dim=(1, 2, 3))
            z = torch.fft.irfftn(y, s=(8, 4096, 4096), dim=(1, 2, 3))
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"check": "cuFFT fp64 rfftn + irfftn of 3 x (8, 4096, 4096), no multiply, no padding/unpadding kernels",
                          "ms": e0.elapsed_time(e1) / 5}), flush=True)
    except Exception as exc:  # noqa: BLE001
        print(json.dumps({"check": "cuFFT via torch unavailable", "error": str(exc)[:200]}), flush=True)

if "c4" in which:
    p = S.Session(lib, cfg("c4", "solvers", n_basis_cells="256 256 1", gneb_n_iterations_amortize=50))
    p.plus_z()
    p.skyrmion(20.0, phase=-90.0)
    p.chain_set_length(64)
    p.jump_to_image(63)
    p.plus_z()
    p.jump_to_image(0)
    p.transition_homogeneous(0, 63)
    p.gneb_start(S.SOLVER_VP, n_iterations=50, n_iterations_log=50)
    n = 500
    t0 = time.time()
    p.gneb_start(S.SOLVER_VP, n_iterations=n, n_iterations_log=n)
    dt = time.time() - t0
    rx, e = p.chain_rx_e()
    print(json.dumps({"config": "C4 GNEB 64 x 256x256x1 VP (through Simulation_GNEB_Start, incl. H2D/D2H of the chain)",
                      "iterations_per_s": n / dt, "image_spin_steps_per_s": 64 * 65536 * n / dt, "barrier_meV": float(e.max() - e[0])}), flush=True)
    p.close()
