#!/usr/bin/env python
"""spin-steps/s of the fused gradient + LLG Depondt step (fp64), BASELINE.json configs[1]:
256^3 simple cubic, exchange + DMI + uniaxial anisotropy, thermal noise T > 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is ONE Depondt iteration over the whole lattice (two fused stage kernels). Prints ONE JSON line (rank 0).
  value        spin-steps/s with the spins resident in HBM (CUDA events on the image's stream, max over ranks)
  e2e          the same metric through the reference-facing C API call Simulation_LLG_Start(Solver_Depondt, n) with
               the spins in HOST memory before and after every call (H2D + D2H inside the timed region)
  roofline     dominant kernel (Depondt stage 2) against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline the reference's own OpenMP implementation (oracle/_ref, built from /root/reference) on a bounded sample
--impl reference times that CPU implementation alone, same metric / unit / config.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from spirit_b200 import capi, session as S, slab  # noqa: E402
from tests import cfgs  # noqa: E402

METRIC = "spin-steps/s (LLG Depondt fp64)"
UNIT = "spin-steps/s"
BYTES_PER_SPIN_STEP = 120.0  # R s | W s' | R s, s' | W s_new, 24 B each, noise regenerated from counters
BYTES_STAGE = (48.0, 72.0)
E2E_BLOCK = 100  # iterations per API call = llg_n_iterations_amortize of the workload
FALLBACK_HBM_GBS = 6650.0


def write_cfg(directory, cells, name="bench.cfg"):
    path = os.path.join(directory, name)
    with open(path, "w") as f:
        f.write(cfgs.render("cubic256", n_basis_cells="%d %d %d" % tuple(cells)))
    return path


def fill_random(sess, seed=20006):
    """uniform on the sphere (z in U[-1,1], phi in U[-pi,pi], like Vectormath.cpp:41-52), written through the live
    spin pointer in chunks"""
    rng = np.random.default_rng(seed)
    sp = sess.spins()
    n, chunk = sp.shape[0], 1 << 21
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        z = rng.uniform(-1, 1, m)
        phi = rng.uniform(-np.pi, np.pi, m)
        r = np.sqrt(1 - z * z)
        sp[i:i + m, 0] = r * np.cos(phi)
        sp[i:i + m, 1] = r * np.sin(phi)
        sp[i:i + m, 2] = z


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(", ") for t, l in self.samples if t0 <= t <= t1] or [l.split(", ") for _, l in self.samples[-3:]]
        sm, reasons, mx, pw = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                pw.append(float(r[2]))
                for name, v in zip(names, r[3:7]):
                    if v.strip().lower() == "active":
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "power_w": float(np.median(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def reference_sample(threads_lib, cells, steps, warmup, tmp):
    """Reference OpenMP Depondt on `cells`; returns (spin-steps/s, seconds, nos)"""
    o = S.Session(threads_lib, write_cfg(tmp, cells, "ref_%d.cfg" % cells[0]))
    fill_random(o)
    if warmup > 0:
        o.llg_start(S.SOLVER_DEPONDT, n_iterations=warmup, n_iterations_log=warmup)
    t0 = time.perf_counter()
    o.llg_start(S.SOLVER_DEPONDT, n_iterations=steps, n_iterations_log=steps)
    dt = time.perf_counter() - t0
    nos = o.nos
    o.close()
    return nos * steps / dt, dt, nos


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    oracle = capi.load_oracle()
    cores = oracle.refshim_num_threads()
    with tempfile.TemporaryDirectory() as tmp:
        # calibrate on 32^3, then choose the largest cubic sample whose K + W steps fit the time budget
        rate, _, _ = reference_sample(oracle, (32, 32, 32), 10, 2, tmp)
        budget = 120.0
        edge = 32
        for cand in (256, 192, 128, 96, 64, 48):
            if cand ** 3 * (args.steps + args.warmup) / rate <= budget:
                edge = cand
                break
        value, dt, nos = reference_sample(oracle, (edge,) * 3, args.steps, args.warmup, tmp)
    sample = "%d^3 sub-lattice of the workload (same Hamiltonian, T, dt), %d Depondt iterations, OMP threads = %d, %s" % (
        edge, args.steps, cores, host_info())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, (edge,) * 3, note="CPU reference on a bounded sample"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, cells, note=None):
    c = {
        "workload": "configs[1]: %dx%dx%d simple cubic, exchange J=10 + DMI D=6 (Bloch) + uniaxial K=1, mu_s=2, "
                    "periodic, LLG Depondt dt=1e-3 alpha=0.3 T=10 K, fp64" % tuple(cells),
        "lattice": list(cells), "solver": "Depondt", "temperature_K": 10.0,
        "l2": "inputs larger than L2 (3 x %.0f MB spin buffers per step vs 126 MB L2)" % (np.prod(cells) * 24 / 1e6),
        "parallelism": "1 GPU" if args.gpus == 1 else (
            "%d GPUs: ONE %dx%dx%d lattice (periodic), slab-decomposed along c, one %dx%dx%d slab per GPU, one-plane halo "
            "exchange per solver stage over NCCL inside the library" % (args.gpus, cells[0], cells[1], cells[2] * args.gpus, cells[0], cells[1], cells[2])),
        "e2e_call": "one Simulation_LLG_Start(Solver_Depondt, n_iterations=steps) call: spins in pinned host memory before, spins + "
                    "effective field in host memory after (H2D 24 B/spin + D2H 48 B/spin inside the timed region), energy/torque "
                    "read back every %d steps" % E2E_BLOCK,
    }
    if note:
        c["note"] = note
    return c


def run_b200(args):
    rank, local_rank, world = dist_env()
    product = capi.load_product()
    if product.SpiritB200_Device_Count() < 1:
        raise SystemExit("bench.py: no CUDA device; spirit_b200 has no CPU fallback")
    product.SpiritB200_Set_Device(local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))
        slab.init_comm(product, dist, rank, world)

    cells = tuple(args.lattice)
    tmp = tempfile.mkdtemp()
    p = S.Session(product, write_cfg(tmp, cells, "bench_%d.cfg" % rank))
    nos = p.nos
    if world > 1:
        # weak scaling: the global lattice has world * Nc planes, this rank owns planes [rank * Nc, (rank + 1) * Nc)
        if product.SpiritB200_Slab_Setup(p.state, rank * cells[2], world * cells[2], -1) != 0:
            raise SystemExit("bench.py: SpiritB200_Slab_Setup failed")
    fill_random(p, seed=20006 + rank)
    p.upload()
    launches0 = p.kernel_launches()

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- device-resident loop: W warm-up steps, then exactly K timed steps ------------------------------------------------
    if args.warmup > 0:
        p.iterate_device(S.SOLVER_DEPONDT, args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = p.kernel_launches()
    t0 = time.perf_counter()
    ms = p.iterate_device(S.SOLVER_DEPONDT, args.steps)
    t1 = time.perf_counter()
    barrier()
    launches = p.kernel_launches() - l0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = nos * world * args.steps / (ms * 1e-3)

    # ---- per-kernel roofline: CUDA events between the stage kernels --------------------------------------------------------
    stage_ms = (capi.ctypes.c_double * 4)()
    n_prof = min(50, args.steps)
    product.SpiritB200_LLG_Profile_Stages(p.state, S.SOLVER_DEPONDT, n_prof, stage_ms, 4, -1)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    fused = p.step_variant(S.SOLVER_DEPONDT) == 2
    if fused:
        # ONE kernel per iteration (sc6_fused.cuh). Algorithmic bytes per launch: SURVEY.md 8d's 120 B per spin-step (the
        # two-pass model the target is quoted on) x the spin-steps one launch processes. The fused kernel's own minimum
        # is 48 B per spin-step (read s, write s_new): `achieved_min_traffic` states the same time against that figure.
        achieved = BYTES_PER_SPIN_STEP * nos / (stage_ms[0] * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": "k_sc6_fused<Depondt,...> (predictor + corrector of one iteration in one launch: gradient(s), "
                                      "noise, virtual force, rotation, s' through shared memory, gradient(s'), rotation)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_SPIN_STEP * nos,
            "model": "SURVEY.md 8d: 120 B per spin-step (R s | W s' | R s, s' | W s_new)",
            "kernel_ms": stage_ms[0],
            "min_traffic": {"bytes_per_spin_step": 48.0, "achieved": 48.0 * nos / (stage_ms[0] * 1e-3) / 1e9,
                            "frac": 48.0 * nos / (stage_ms[0] * 1e-3) / 1e9 / peak,
                            "note": "the fused kernel reads s once and writes s_new once; s' never leaves the SM"},
            "step": {"bytes_per_spin_step": BYTES_PER_SPIN_STEP},
        }
    else:
        k = 1  # stage 2 moves 72 of the 120 B
        achieved = BYTES_STAGE[k] * nos / (stage_ms[k] * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": "k_sc6_stage<Depondt,2,...> (gradient(s) recomputed + gradient(s') + virtual forces + Rodrigues rotation)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_STAGE[k] * nos,
            "stage_ms": [stage_ms[0], stage_ms[1]],
            "stage1": {"achieved": BYTES_STAGE[0] * nos / (stage_ms[0] * 1e-3) / 1e9, "bytes_per_spin": BYTES_STAGE[0]},
            "step": {"bytes_per_spin_step": BYTES_PER_SPIN_STEP},
        }
    roofline["step"]["achieved"] = BYTES_PER_SPIN_STEP * (value / world) / 1e9
    roofline["step"]["frac"] = roofline["step"]["achieved"] / peak
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            roofline["traffic"] = json.load(open(traffic_path)).get(
                "k_sc6_fused_depondt_bytes_per_launch_256" if fused else "k_sc6_stage_depondt_2_bytes_per_launch_256")
            if cells != (256, 256, 256):
                roofline["traffic"] = None
        except (ValueError, OSError):
            pass

    # ---- end to end through the C API with host buffers ---------------------------------------------------------------------
    # ONE reference-facing call for the K timed steps: Simulation_LLG_Start(Solver_Depondt, n_iterations=K). Before the call
    # the spins are in (pinned) host memory behind System_Get_Spin_Directions, after it spins AND effective field are back
    # in host memory; every llg_n_iterations_amortize (= 100) steps the hook reads energy and max torque back to the host.
    e2e = None
    if not args.no_e2e:
        p.llg_start(S.SOLVER_DEPONDT, n_iterations=E2E_BLOCK, n_iterations_log=E2E_BLOCK)  # warm-up call
        barrier()
        l1 = p.kernel_launches()
        te0 = time.perf_counter()
        p.llg_start(S.SOLVER_DEPONDT, n_iterations=args.steps, n_iterations_log=args.steps)
        _ = float(p.energy())  # the run's result on the host
        te = time.perf_counter() - te0
        e2e_launches = p.kernel_launches() - l1
        if dist is not None:
            import torch
            t = torch.tensor([te], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        n_hooks = max(1, args.steps // E2E_BLOCK)
        e2e = {"value": nos * world * args.steps / te, "unit": UNIT,
               "h2d_bytes_per_step": 24.0 * nos / args.steps, "d2h_bytes_per_step": (48.0 * nos + 16.0 * n_hooks) / args.steps,
               "calls": 1, "iterations_per_call": args.steps, "seconds": te, "gpu_launches": int(e2e_launches)}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference OpenMP build on a bounded sample ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        oracle = capi.load_oracle()
        cores = oracle.refshim_num_threads()
        rate, _, _ = reference_sample(oracle, (32, 32, 32), 10, 2, tmp)
        edge = 128 if 128 ** 3 * 12 / rate <= 25.0 else 64
        steps = max(4, min(200, int(15.0 * rate / edge ** 3)))
        v, dt, n = reference_sample(oracle, (edge,) * 3, steps, 1, tmp)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": "%d^3 sub-lattice, %d Depondt iterations in %.1f s, OMP threads = %d, %s" % (edge, steps, dt, cores, host_info())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, cells), "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    p.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lattice", type=int, nargs=3, default=[256, 256, 256], help="debug: override the lattice")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
