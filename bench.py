#!/usr/bin/env python
"""spin-steps/s of the fused gradient + LLG Depondt step (fp64), BASELINE.json configs[1]:
256^3 simple cubic, exchange + DMI + uniaxial anisotropy, thermal noise T > 0.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extras]

A "step" is ONE Depondt iteration over the whole lattice: one launch of the fused predictor + corrector kernel
(spirit_b200/csrc/device/sc6_fused.cuh). Prints ONE JSON line (rank 0).
  value         spin-steps/s with the spins resident in HBM (CUDA events on the image's stream, max over ranks)
  e2e           the same metric through the reference-facing C API call Simulation_LLG_Start(Solver_Depondt, n) with
                the spins in HOST memory before and after every call (H2D + D2H inside the timed region)
  roofline      the step kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own OpenMP implementation (oracle/_ref, built from /root/reference) on a bounded sample
  configs       the other BASELINE configurations, each a bounded run in the same process (skipped with --no-extras):
                c1 (100x100x1 default input.cfg), c3 (2048x2048x4 film + dipolar FFT convolution, cuFFT timed beside it as a
                check), c4 (GNEB skyrmion collapse, 64 images of 256x256 with a climbing image), c5 (512^3 + dipolar
                convolution, SIB, STRONG scaling: the same lattice on every N)
  multi_gpu_parity  (N > 1) slabs with the in-kernel halo exchange and the distributed dipolar convolution against the same
                lattices on rank 0's GPU alone: maximum deviation of the spins after a few iterations
--impl reference times the reference's CPU implementation alone, same metric / unit / config.
"""
import argparse
import ctypes
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
BYTES_FUSED = 48.0           # what the fused kernel has to move: R s | W s_new
BYTES_STAGE = (48.0, 72.0)
E2E_BLOCK = 100  # iterations per API call = llg_n_iterations_amortize of the workload
FALLBACK_HBM_GBS = 6650.0


def write_cfg(directory, cells, name="bench.cfg", preset="cubic256", **over):
    path = os.path.join(directory, name)
    with open(path, "w") as f:
        f.write(cfgs.render(preset, n_basis_cells="%d %d %d" % tuple(cells), **over))
    return path


def unit_random(n, seed):
    """uniform on the sphere (z in U[-1,1], phi in U[-pi,pi], like Vectormath.cpp:41-52)"""
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1, 1, n)
    phi = rng.uniform(-np.pi, np.pi, n)
    r = np.sqrt(1 - z * z)
    return np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)


def fill_random(sess, seed=20006):
    """the same, written through the live spin pointer in chunks"""
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
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
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


# ---- the reference's CPU implementation (oracle/_ref): cpu_baseline leg and the --impl reference arm --------------------------
def reference_threads(oracle):
    """All the host threads this process may use. torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the
    reference arm must not inherit that (one thread would be a crippled baseline), so the OpenMP thread count is set
    explicitly to the cores of the process' affinity mask."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    oracle.refshim_set_num_threads(cores)
    return oracle.refshim_num_threads()


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


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    oracle = capi.load_oracle()
    cores = reference_threads(oracle)
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
        "config": workload_config(args, tuple(args.lattice)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, cells):
    return {
        "workload": "configs[1]: %dx%dx%d simple cubic, exchange J=10 + DMI D=6 (Bloch) + uniaxial K=1, mu_s=2, "
                    "periodic, LLG Depondt dt=1e-3 alpha=0.3 T=10 K, fp64" % tuple(cells),
        "lattice": list(cells), "solver": "Depondt", "temperature_K": 10.0,
        "l2": "inputs larger than L2 (2 x %.0f MB spin buffers per step vs 126 MB L2)" % (np.prod(cells) * 24 / 1e6),
        "parallelism": "1 GPU" if args.gpus == 1 else (
            "%d GPUs: ONE %dx%dx%d lattice (periodic), slab-decomposed along c, one %dx%dx%d slab per GPU; the CTAs at the slab "
            "ends store their two outermost planes into the neighbours' halo planes over NVLink (peer-mapped memory), ranks keep "
            "in step with stream memory operations: no collective on the data path" % (
                args.gpus, cells[0], cells[1], cells[2] * args.gpus, cells[0], cells[1], cells[2])),
        "e2e_call": "one Simulation_LLG_Start(Solver_Depondt, n_iterations=steps) call: spins in pinned host memory before and after "
                    "(H2D 24 B/spin + D2H 24 B/spin inside the timed region), energy/torque read back every %d steps; the effective "
                    "field of the last hook stays on the device and is mirrored when System_Get_Effective_Field asks for it" % E2E_BLOCK,
    }


# ---- the other BASELINE configurations (bounded runs; each returns a dict, never raises) -----------------------------------
def guarded(fn):
    def run(*a, **kw):
        t0 = time.perf_counter()
        try:
            out = fn(*a, **kw)
        except Exception as exc:  # noqa: BLE001
            out = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
        out["wall_s"] = round(time.perf_counter() - t0, 2)
        return out
    return run


@guarded
def config_c1(lib, tmp, peak):
    """configs[0]: the reference's default input.cfg (100x100x1, J + DMI + field, Depondt, T = 0): launch-latency regime"""
    p = S.Session(lib, write_cfg(tmp, (100, 100, 1), "c1.cfg", preset="default"))
    p.plus_z()
    p.skyrmion(5.0, phase=-90.0)
    p.upload()
    p.iterate_device(S.SOLVER_DEPONDT, 200)
    n = 5000
    ms = p.iterate_device(S.SOLVER_DEPONDT, n)
    out = {"workload": "configs[0]: 100x100x1, default input.cfg, Depondt T=0", "iterations_per_s": n / ms * 1e3,
           "spin_steps_per_s": p.nos * n / ms * 1e3, "us_per_iteration": ms / n * 1e3, "step_variant": p.step_variant(S.SOLVER_DEPONDT)}
    p.close()
    return out


def cufft_check(shape, reps=5):
    """cuFFT (through torch.fft) timed alongside as a CHECK only: the library-style, un-pruned execution of the transforms of one
    dipolar convolution (3 forward R2C + 3 inverse C2R of the padded lattice; multiply, padding and un-padding not included)"""
    try:
        import torch
        x = torch.zeros((3,) + tuple(shape), dtype=torch.float64, device="cuda")
        x[:, :shape[0] // 2, :shape[1] // 2, :shape[2] // 2] = 1.0
        for _ in range(2):
            z = torch.fft.irfftn(torch.fft.rfftn(x, dim=(1, 2, 3)), s=tuple(shape), dim=(1, 2, 3))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            z = torch.fft.irfftn(torch.fft.rfftn(x, dim=(1, 2, 3)), s=tuple(shape), dim=(1, 2, 3))
        e1.record()
        torch.cuda.synchronize()
        del x, z
        torch.cuda.empty_cache()
        return {"what": "cuFFT fp64 rfftn + irfftn of 3 x %s via torch.fft: transforms only, a check, not part of the product" % (tuple(shape),),
                "ms": e0.elapsed_time(e1) / reps}
    except Exception as exc:  # noqa: BLE001
        return {"what": "cuFFT via torch unavailable", "error": str(exc)[:200]}


@guarded
def config_c3(lib, tmp, peak):
    """configs[2]: 2048x2048x4 open film, exchange + DMI + dipolar FFT convolution; VP minimiser and LLG Depondt"""
    p = S.Session(lib, write_cfg(tmp, (2048, 2048, 4), "c3.cfg", boundary_conditions="0 0 0", ddi_method="fft",
                                 ddi_n_periodic_images="0 0 0", external_field_magnitude=25, anisotropy_magnitude=0, llg_temperature=0))
    fill_random(p)
    p.upload()  # builds the plan: tensor + its spectrum
    nos = p.nos
    out = {"workload": "configs[2]: 2048x2048x4 film, open, exchange + DMI + dipolar FFT convolution (padded 4096x4096x8)"}
    for solver, name, n_eval, other in ((S.SOLVER_VP, "VP", 1, 144.0), (S.SOLVER_DEPONDT, "Depondt", 2, 120.0)):
        p.iterate_device(solver, 3)
        n = 20
        ms = p.iterate_device(solver, n) / n
        out[name] = {
            "ms_per_iteration": ms, "spin_steps_per_s": nos / ms * 1e3,
            "ddi_evaluations_per_iteration": n_eval,
            # SURVEY.md 8d model of one pruned convolution: 1008 B per spin with a complex tensor spectrum, 816 B with the real one
            # the shipped kernels read (single sublattice); + the stencil / solver bytes of the iteration
            "roofline_1008": {"bytes_per_spin_step": other + n_eval * 1008.0, "achieved": (other + n_eval * 1008.0) * nos / ms / 1e6,
                              "frac": (other + n_eval * 1008.0) * nos / ms / 1e6 / peak},
            "roofline_816": {"bytes_per_spin_step": other + n_eval * 816.0, "achieved": (other + n_eval * 816.0) * nos / ms / 1e6,
                             "frac": (other + n_eval * 816.0) * nos / ms / 1e6 / peak},
        }
    p.close()
    out["check"] = cufft_check((8, 4096, 4096))
    return out


@guarded
def config_c4(lib, tmp, peak, dist=None, rank=0, world=1):
    """configs[3]: GNEB skyrmion collapse, 64 images of 256x256x1 (the reference's solvers.cfg physics: the barrier of its own
    GNEB test, core/test/test_solvers.cpp:74-103), climbing image, VP. N > 1: whole images per GPU (64 / N consecutive images
    each, SpiritB200_Chain_Shard_Setup): boundary images to the neighbouring ranks and the per-image scalars shared per force
    evaluation; the same chain, so the barrier must be the N = 1 one."""
    noi = 64
    p = S.Session(lib, write_cfg(tmp, (256, 256, 1), "c4_%d.cfg" % rank, preset="solvers", gneb_n_iterations_amortize=50,
                                 llg_n_iterations_amortize=100, llg_force_convergence="1e-7"))
    p.plus_z()
    p.skyrmion(5.0, phase=-90.0)
    p.llg_set(direct_minimization=True)
    p.llg_start(S.SOLVER_VP, n_iterations=20000, n_iterations_log=20000)  # relax the metastable skyrmion of image 0
    p.chain_set_length(noi)
    p.jump_to_image(noi - 1)
    p.plus_z()
    p.jump_to_image(0)
    p.transition_homogeneous(0, noi - 1)
    i_begin, n_local = 0, noi
    if world > 1:
        # every rank built the same initial chain on the host; it keeps its consecutive images
        from spirit_b200 import slab
        i_begin, n_local = slab.partition(noi, world)[rank]
        images0 = [p.spins(i).copy() for i in range(i_begin, i_begin + n_local)]
        p.close()
        p = S.Session(lib, write_cfg(tmp, (256, 256, 1), "c4s_%d.cfg" % rank, preset="solvers", gneb_n_iterations_amortize=50))
        p.chain_set_length(n_local)
        for i in range(n_local):
            p.set_spins(images0[i], idx_image=i)
        if lib.SpiritB200_Chain_Shard_Setup(p.state, i_begin, noi) != 0:
            raise RuntimeError("SpiritB200_Chain_Shard_Setup failed")

    def energies():
        rx, e = p.chain_rx_e()
        if world == 1:
            return np.asarray(e)
        parts = [None] * world
        dist.all_gather_object(parts, np.asarray(e))
        return np.concatenate(parts)

    p.gneb_start(S.SOLVER_VP, n_iterations=3000, n_iterations_log=3000)
    if world == 1:
        p.gneb_set_image_type_automatically()
    else:  # Parameters_GNEB_Set_Image_Type_Automatically over the whole chain (maxima climb, minima fall)
        e = energies()
        for g in range(max(1, i_begin), min(noi - 1, i_begin + n_local)):
            if e[g - 1] < e[g] > e[g + 1]:
                p.gneb_set_image_type(S.GNEB_CLIMBING, g - i_begin)
            elif e[g - 1] > e[g] < e[g + 1]:
                p.gneb_set_image_type(S.GNEB_FALLING, g - i_begin)
    p.gneb_start(S.SOLVER_VP, n_iterations=3000, n_iterations_log=3000)
    n = 1000
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    p.gneb_start(S.SOLVER_VP, n_iterations=n, n_iterations_log=n)
    dt = time.perf_counter() - t0
    tq = float(p.chain_max_torque())
    if dist is not None:
        import torch
        t = torch.tensor([dt, tq], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, tq = float(t[0].item()), float(t[1].item())
    e = energies()
    k = int(np.argmax(e))
    out = {"workload": "configs[3]: GNEB skyrmion collapse, 64 images of 256x256x1, climbing image (set automatically), VP; timed "
                       "through Simulation_GNEB_Start incl. H2D / D2H of the chain; %s" % (
                           "one GPU" if world == 1 else "%d images per GPU on %d GPUs (max over ranks)" % (noi // world, world)),
           "n_gpus": world, "iterations_per_s": n / dt, "image_spin_steps_per_s": noi * 65536 * n / dt, "barrier_meV": float(e[k] - e[0]),
           "saddle_image": k, "max_torque": tq, "iterations_before_timing": 6000}
    p.close()
    return out


@guarded
def config_c5(lib, tmp, peak, dist, rank, world, steps=5):
    """configs[4]: 512^3 simple cubic, exchange + DMI + dipolar FFT convolution, LLG SIB; STRONG scaling: the same lattice on every
    N (slabs of 512 / N planes, distributed convolution with the kb axis cut over the ranks)"""
    N, ncl = 512, 512 // world
    p = S.Session(lib, write_cfg(tmp, (N, N, ncl), "c5_%d.cfg" % rank, boundary_conditions="0 0 0", ddi_method="fft",
                                 ddi_n_periodic_images="0 0 0", anisotropy_magnitude=0, external_field_magnitude=25, llg_temperature=0))
    if world > 1 and lib.SpiritB200_Slab_Setup(p.state, rank * ncl, N, -1) != 0:
        raise RuntimeError("SpiritB200_Slab_Setup failed")
    fill_random(p, seed=7 + rank)
    t0 = time.perf_counter()
    p.upload()
    setup = time.perf_counter() - t0
    p.iterate_device(S.SOLVER_SIB, 2)
    if dist is not None:
        dist.barrier()
    ms = p.iterate_device(S.SOLVER_SIB, steps)
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    p.close()
    nos = N ** 3
    return {"workload": "configs[4]: 512^3 sc, exchange + DMI + dipolar FFT convolution (padded 1024^3), LLG SIB, %d GPU(s), slabs of "
                        "%d planes" % (world, ncl), "scaling": "strong", "n_gpus": world, "iterations": steps,
            "ms_per_iteration": ms / steps, "spin_steps_per_s": nos * steps / ms * 1e3, "ddi_plan_setup_s": setup,
            "model_bytes_per_spin_step": 120 + 2 * 816, "per_gpu_model_GBps": nos * steps / ms * 1e3 * (120 + 2 * 816) / 1e9 / world}


@guarded
def multi_gpu_parity(lib, tmp, dist, rank, world):
    """Correctness carried on the same line as the speed: two reduced lattices evolved on `world` slabs and on rank 0's GPU alone.
    (1) the bench Hamiltonian at T > 0, Depondt (fused kernel, in-kernel halo exchange): bit-identical by construction;
    (2) exchange + DMI + dipolar convolution, SIB (distributed transposes)."""
    out = {"n_gpus": world}
    cases = (("slabs_depondt_T10", "cubic256", (64, 48, 16 * world), dict(llg_temperature=10, boundary_conditions="1 1 1"), S.SOLVER_DEPONDT, 8),
             ("distributed_ddi_sib", "default", (64, 32, 8 * world), dict(boundary_conditions="0 0 0", ddi_method="fft",
                                                                         ddi_n_periodic_images="0 0 0"), S.SOLVER_SIB, 4))
    for name, preset, (Na, Nb, Nc), over, solver, n in cases:
        s0 = unit_random(Na * Nb * Nc, 99)
        ncl = Nc // world
        p = S.Session(lib, write_cfg(tmp, (Na, Nb, ncl), "par_%s_%d.cfg" % (name, rank), preset=preset, llg_n_iterations_amortize=4, **over))
        if lib.SpiritB200_Slab_Setup(p.state, rank * ncl, Nc, -1) != 0:
            raise RuntimeError("SpiritB200_Slab_Setup failed")
        p.set_spins(s0[rank * ncl * Na * Nb:(rank + 1) * ncl * Na * Nb])
        p.llg_start(solver, n_iterations=n, n_iterations_log=n)
        mine, e_slab = p.spins().copy(), p.energy()
        p.close()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            g = S.Session(lib, write_cfg(tmp, (Na, Nb, Nc), "par_%s_global.cfg" % name, preset=preset, llg_n_iterations_amortize=4, **over))
            g.set_spins(s0)
            g.llg_start(solver, n_iterations=n, n_iterations_log=n)
            ref = g.spins()
            out[name] = {"lattice": [Na, Nb, Nc], "iterations": n, "max_spin_deviation": float(np.abs(np.concatenate(parts) - ref).max()),
                         "moved": float(np.abs(ref - s0).max()), "energy_rel_deviation": float(abs(e_slab - g.energy()) / abs(g.energy()))}
            g.close()
        dist.barrier()
    return out


def thermal_fp64_check(product_ms):
    """The headline step of a build whose only difference is an fp64 Box-Muller (log, sqrt, sincospi) for the thermal variates:
    what the fp32 / SFU shaping of the product buys. A check, not part of the product path."""
    lib = "libSpirit_xi64.so"
    if not os.path.exists(os.path.join(ROOT, "spirit_b200", lib)):
        return {"unavailable": "spirit_b200/%s not built (__graft_entry__.build())" % lib}
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "50", "--warmup", "5", "--no-e2e",
                            "--no-cpu-baseline", "--no-extras"], env=dict(os.environ, SPIRIT_B200_LIB=lib), capture_output=True,
                           text=True, timeout=300)
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        return {"what": "same step, thermal variates shaped in fp64 (-DSB_THERMAL_FP64=1)", "ms_per_step": d["ms_per_step"],
                "product_ms_per_step": product_ms, "slowdown": d["ms_per_step"] / product_ms}
    except Exception as exc:  # noqa: BLE001
        return {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}


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

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident loop: W warm-up steps, then exactly K timed steps ------------------------------------------------
    # (the clock sampler starts BEFORE the warm-up: nothing may idle the GPU between the warm-up and the timed region except the
    # barrier + synchronize the contract asks for; a 0.3 s pause there let the clocks drop and put their ramp inside the 9 ms region)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    p.iterate_device(S.SOLVER_DEPONDT, 200)  # 90 ms of the same loop: clocks at their load level before the W warm-up steps
    if args.warmup > 0:
        p.iterate_device(S.SOLVER_DEPONDT, args.warmup)  # returns after a device synchronize
    barrier()
    l0 = p.kernel_launches()
    t0 = time.perf_counter()
    ms = p.iterate_device(S.SOLVER_DEPONDT, args.steps)
    t1 = time.perf_counter()
    barrier()
    launches = p.kernel_launches() - l0
    # K steps of this workload last a few milliseconds, nvidia-smi samples every 50 ms: the same loop keeps running (untimed) right
    # behind the timed region until the sampler has seen the GPU under this load for about a second; the clocks line reports the
    # samples from the start of the timed region to the end of that tail
    t_load = time.perf_counter()
    for _ in range(25):  # (a fixed count: slabs must run the same number of iterations on every rank)
        p.iterate_device(S.SOLVER_DEPONDT, 100)
    t2 = time.perf_counter()
    barrier()
    clocks = sampler.stop(t0, t2) if rank == 0 else None
    if clocks is not None:
        clocks["sampled"] = "timed region (%.1f ms) + %.1f s of the same loop, untimed, directly behind it; nvidia-smi every 50 ms" % (
            (t1 - t0) * 1e3, t2 - t_load)
    ms = max_over_ranks(ms)
    value = nos * world * args.steps / (ms * 1e-3)

    # ---- per-kernel roofline: CUDA events between the kernels of an iteration ------------------------------------------------
    stage_ms = (ctypes.c_double * 4)()
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
        # two-pass model the target is quoted on) x the spin-steps one launch processes. The fused kernel itself has to move
        # only 48 B per spin-step (read s, write s_new; the predictor never leaves the SM): `min_traffic` states the same
        # time against that figure, and `traffic` (ncu) is what it does move.
        achieved = BYTES_PER_SPIN_STEP * nos / (stage_ms[0] * 1e-3) / 1e9
        roofline = {
            "bound": "hbm", "kernel": "k_sc6_fused<Depondt,...> (predictor + corrector of one iteration in one launch: gradient(s), "
                                      "noise, virtual force, rotation, s' through shared memory, gradient(s'), rotation)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_SPIN_STEP * nos,
            "model": "SURVEY.md 8d: 120 B per spin-step (R s | W s' | R s, s' | W s_new)",
            "kernel_ms": stage_ms[0],
            "min_traffic": {"bytes_per_spin_step": BYTES_FUSED, "achieved": BYTES_FUSED * nos / (stage_ms[0] * 1e-3) / 1e9,
                            "frac": BYTES_FUSED * nos / (stage_ms[0] * 1e-3) / 1e9 / peak,
                            "note": "the fused kernel reads s once and writes s_new once; it is bound by instruction issue and "
                                    "fp64 latency, not by HBM (DESIGN.md 3.2)"},
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
    if os.path.exists(traffic_path) and cells == (256, 256, 256):
        try:
            roofline["traffic"] = json.load(open(traffic_path)).get(
                "k_sc6_fused_depondt_bytes_per_launch_256" if fused else "k_sc6_stage_depondt_2_bytes_per_launch_256")
        except (ValueError, OSError):
            pass

    # ---- end to end through the C API with host buffers ---------------------------------------------------------------------
    # ONE reference-facing call for the K timed steps: Simulation_LLG_Start(Solver_Depondt, n_iterations=K). Before the call
    # the spins are in (pinned) host memory behind System_Get_Spin_Directions, after it the new spins are back there; every
    # llg_n_iterations_amortize (= 100) steps the hook reads energy and max torque back to the host. (The effective field is
    # mirrored lazily, on System_Get_Effective_Field: not part of this call.)
    e2e = None
    if not args.no_e2e:
        p.llg_start(S.SOLVER_DEPONDT, n_iterations=E2E_BLOCK, n_iterations_log=E2E_BLOCK)  # warm-up call
        barrier()
        l1 = p.kernel_launches()
        te0 = time.perf_counter()
        p.llg_start(S.SOLVER_DEPONDT, n_iterations=args.steps, n_iterations_log=args.steps)
        _ = float(p.energy())  # the run's result on the host
        te = max_over_ranks(time.perf_counter() - te0)
        e2e_launches = p.kernel_launches() - l1
        n_hooks = max(1, args.steps // E2E_BLOCK)
        e2e = {"value": nos * world * args.steps / te, "unit": UNIT,
               "h2d_bytes_per_step": 24.0 * nos / args.steps, "d2h_bytes_per_step": (24.0 * nos + 16.0 * n_hooks) / args.steps,
               "calls": 1, "iterations_per_call": args.steps, "seconds": te, "gpu_launches": int(e2e_launches)}
    p.close()

    # ---- the other BASELINE configurations and the multi-GPU correctness record ---------------------------------------------------
    configs, parity = None, None
    if not args.no_extras:
        configs = {}
        if world == 1:
            configs["c1"] = config_c1(product, tmp, peak)
            configs["c3"] = config_c3(product, tmp, peak)
            configs["c4"] = config_c4(product, tmp, peak)
        else:
            # (the sharded chain is timed BEFORE the parity record: behind it the same 1000 iterations took 1.4 to 5 times longer
            # in two runs on 2 GPUs, profiles/r2zm, r2zo -- something the slab / distributed-convolution sessions of the parity
            # record leave behind costs the chain's ncclSend/Recv + all-reduces; the kernels are the same)
            if 64 % world == 0:
                configs["c4"] = config_c4(product, tmp, peak, dist, rank, world)
        configs["c5"] = config_c5(product, tmp, peak, dist, rank, world)
        if world > 1:
            parity = multi_gpu_parity(product, tmp, dist, rank, world)

    # ---- check: the same step with the thermal variates shaped in fp64 (build variant libSpirit_xi64.so, own process) ----------
    checks = None
    if rank == 0 and world == 1 and not args.no_extras:
        checks = {"thermal_fp64": thermal_fp64_check(ms / args.steps)}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the reference OpenMP build on a bounded sample ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        oracle = capi.load_oracle()
        cores = reference_threads(oracle)
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
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "configs": configs,
            "multi_gpu_parity": parity, "checks": checks,
        }
        print(json.dumps(line), flush=True)
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
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations and the multi-GPU parity record")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
