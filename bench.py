#!/usr/bin/env python
"""bench.py -- chimera-b200 headline benchmark.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on and the
largest that fits one GPU): uniform thermal plasma, Nx=4096, Nr=512, M=1 (modes 0,1),
16 particles per cell, electrons + immobile ions, DampCells=50, no laser, no frame.
A "step" is one full PIC_loop.step() (push+sort+deposit+transforms+PSATD+gather).
Metric: particle-steps/s = mobile particles x steps / time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1: launched by torchrun, one rank per GPU; every rank holds the full grid and its
own 16 ppc shard of particles (weak scaling: per-GPU particle work fixed), rho/J are
summed with NCCL every step, and the field solve on the (same-sized) grid is split over
the ranks by kr rows (Solver.enable_spectral_sharding; --replicated-solve: every rank
solves the whole grid, as in the single-GPU run).  --impl reference times the reference's own CPU implementation
(oracle/_ref = chimeraCL's kernels host-compiled + OpenMP, np.dot, np.fft).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# bytes per mobile particle-step, SURVEY.md 8(d): reads+writes each stage must move
BYTES = {"push": 80, "sort": 40, "depose_vector": 68, "depose_scalar": 36, "gather": 84}
BYTES_PER_PARTICLE_STEP = 2 * BYTES["push"] + 2 * BYTES["sort"] + BYTES["depose_vector"] \
    + BYTES["depose_scalar"] + BYTES["gather"]          # 428


def workload(small=False):
    if small:       # debugging only (not a bench configuration)
        Nx, Nr = 512, 128
    else:
        Nx, Nr = 4096, 512
    half = 0.025 * Nx
    grid = {"Xmin": -half, "Xmax": half, "Nx": Nx, "Rmin": 0.0, "Rmax": 0.0125 * Nx,
            "Nr": Nr, "M": 1, "DampCells": 50}
    grid["dt"] = (grid["Xmax"] - grid["Xmin"]) / grid["Nx"]
    return grid


def species_cfgs(solver_args, nppc=(2, 2, 4)):
    eons = {"Nppc": nppc, "dx": solver_args["dx"], "dr": solver_args["dr"],
            "dt": solver_args["dt"], "dens": 0.01, "charge": -1}
    ions = dict(eons, charge=1, Immobile=True)
    return eons, ions


def plasma_domain(A):
    """All valid cells: ix in [1, Nx-3], ir in [0, Nr-3]."""
    dx, dr = A["dx"], A["dr"]
    return {"Xmin": A["Xmin"] + dx, "Xmax": A["Xmin"] + dx + (A["Nx"] - 3 - 0.5) * dx,
            "Rmin": 0.0, "Rmax": (A["Nr"] - 2) * dr,
            "dpx": 0.01, "dpy": 0.01, "dpz": 0.01}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ============================================================================ reference arm
def run_reference(args, as_baseline=False):
    """chimeraCL's own CPU path on the host cores: its OpenCL C kernels compiled for
    the host (oracle/_ref, OpenMP over work-items) + np.dot + np.fft, driven by the
    restated methods/ orchestration.  Bounded sample of the workload: full grid,
    particles subsampled to 2 ppc; the particle phases are scaled back to 16 ppc
    (they are linear in the particle count), the field phases are measured as is."""
    from oracle import orchestration as O
    from oracle.ref_kernels import RefKernels, ref_available
    from oracle.np_kernels import NumpyKernels
    cores = os.cpu_count() or 1
    M = 1
    if ref_available(M):
        K = RefKernels(M, parallel=True)
        kind = "reference"
    else:
        K = NumpyKernels(M)
        kind = "port"
    grid = workload(args.small)
    S = O.OracleSolver(dict(grid), K)
    A = S.Args
    full_nppc, sample_nppc = (2, 2, 4), (1, 1, 2)
    scale = int(np.prod(full_nppc) // np.prod(sample_nppc))
    ecfg, icfg = species_cfgs(A, full_nppc)
    dom = plasma_domain(A)
    rng = np.random.default_rng(1234)
    Nx_loc = int(np.ceil((dom["Xmax"] - dom["Xmin"]) / A["dx"]) + 1)
    Nr_loc = int(np.round((dom["Rmax"] - dom["Rmin"]) / A["dr"]) + 1)
    xg = dom["Xmin"] + A["dx"] * np.arange(Nx_loc)
    rg = dom["Rmin"] + A["dr"] * np.arange(Nr_loc)
    th = rng.uniform(0, 2 * np.pi, (Nx_loc - 1) * (Nr_loc - 1))
    x, y, z, w = NumpyKernels(M).fill_grid(th, xg, rg, sample_nppc)
    n = x.size
    P = O.OracleParticles(ecfg, K)
    w = w * P.Args["w0"]
    px, py, pz = (rng.normal(0, 0.01, n) for _ in range(3))
    P.set_particles(x=x, y=y, z=z, px=px, py=py, pz=pz, w=w,
                    g_inv=1 / np.sqrt(1 + px * px + py * py + pz * pz))
    I = O.OracleParticles(icfg, K)
    I.set_particles(x=x.copy(), y=y.copy(), z=z.copy(), w=w.copy())
    species = [P, I]
    np_full = n * scale

    def particle_phases_a():
        for p in species:
            p.push_coords("half")
            p.sort_parts(S)
        S.depose_currents(species)
        for p in species:
            p.push_coords("half")
            p.sort_parts(S)
        S.depose_charge(species)

    def field_phases():
        S.fb_transform(scals=["rho"], vects=["J"], dir=0)
        S.fields_smooth(["rho", "Jx", "Jy", "Jz"])
        for m in range(S.M + 1):
            for c in "xyz":
                S.D["dN0%s_fb_m%d" % (c, m)][...] = S.D["dN1%s_fb_m%d" % (c, m)]
        S.field_grad("rho", "dN1")
        S.push_fields()
        S.damp_fields()
        S.restore_B_fb()
        S.fb_transform(vects=["E", "B"], dir=1)

    def one_step(with_fields):
        t0 = time.perf_counter()
        particle_phases_a()
        t1 = time.perf_counter()
        if with_fields:
            field_phases()
        t2 = time.perf_counter()
        S.gather_and_push(species)
        t3 = time.perf_counter()
        return (t1 - t0) + (t3 - t2), (t2 - t1)

    # The field phases cost the same every step (data independent): they are timed
    # in the first timed step only and that time is charged to every step, so that
    # the run stays within minutes on a few host cores.
    steps = max(1, args.steps if not as_baseline else 2)
    warm = args.warmup if not as_baseline else 1
    for i in range(warm):
        one_step(with_fields=(i == 0))
    tp = tf = 0.0
    for i in range(steps):
        a, b = one_step(with_fields=(i == 0))
        tp += a
        if i == 0:
            tf = b * steps
    t_step_full = (tp * scale + tf) / steps          # extrapolated 16 ppc step
    value = np_full / t_step_full
    sample = ("full grid Nx=%d Nr=%d M=1; particles subsampled to %d ppc (%d mobile + as many "
              "ions), particle phases scaled x%d to 16 ppc, field phases as measured; "
              "%d steps (field phases timed in the first one); particle %.2f s/step, fields %.2f s/step" %
              (A["Nx"], A["Nr"], int(np.prod(sample_nppc)), n, scale, steps,
               tp * scale / steps, tf / steps))
    cpu = {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind,
           "sample": sample}
    if as_baseline:
        return cpu
    line = {"impl": "reference", "metric": "particle-steps/s", "value": value,
            "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": t_step_full * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(A, np_full, args.gpus),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit(line)


def bench_config(A, np_per_gpu, n_gpus):
    return {"workload": "uniform thermal plasma (BASELINE configs[2])", "Nx": int(A["Nx"]),
            "Nr": int(A["Nr"]), "modes": "m=0,1", "ppc": 16,
            "mobile_particles_per_gpu": int(np_per_gpu),
            "immobile_particles_per_gpu": int(np_per_gpu),
            "parallelism": "particles sharded x%d, grid replicated, NCCL all-reduce of rho/J"
                           % n_gpus,
            "l2": "inputs exceed L2 (2.1 GB of particle data, 2.2 GB of fields per step)"}


# ============================================================================ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from chimeracl_b200 import _lib
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    from chimeracl_b200.pic_loop import PIC_loop
    from chimeracl_b200.parallel import init_distributed
    from chimeracl_b200 import host_api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm = Communicator(answers=[0, 0], seed=1234 + rank)
    init_distributed(comm)
    dev = comm.device

    grid = workload(args.small)
    solver = Solver(dict(grid), comm)
    A = solver.Args
    ecfg, icfg = species_cfgs(A)
    eons = Particles(ecfg, comm)
    ions = Particles(icfg, comm)
    ions.Args["InjectorSource"] = eons
    eons.make_new_domain(plasma_domain(A))
    eons.add_new_particles()
    ions.add_new_particles(source=eons)
    eons.free_added()
    for p in (eons, ions):
        p.sort_parts(solver)
        p.align_parts()
    np_gpu = int(eons.Args["Np"])
    # N > 1: kr-row sharded field solve (parity against the replicated
    # solve 1.3e-14 over NCCL on 2 and 4 B200, 7.24 -> 6.17 and 7.40 -> 5.61 ms/step);
    # --replicated-solve or CHB_SHARD_SPECTRAL=0 brings the replicated solve back
    shard_solve = world > 1 and not args.replicated_solve and \
        os.environ.get("CHB_SHARD_SPECTRAL", "1") != "0"
    if shard_solve:
        solver.enable_spectral_sharding()
    loop = PIC_loop(solvers=[solver], species=[eons, ions])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        loop.step()
    sampler = ClockSampler(dev.index)
    sampler.start()
    _lib.CALL_COUNTS.clear()
    ms = timed(loop.step, args.steps)
    launches = _lib.launches()
    clocks = sampler.stop()
    ms_step = ms / args.steps
    value = np_gpu * world * args.steps / (ms * 1e-3)

    # ---- phase split and per-kernel times (CUDA events around every C-ABI call)
    loop.timit = True
    loop.Timer = {k: 0 for k in __import__("chimeracl_b200.pic_loop", fromlist=["x"]).loop_steps}
    loop._events = []
    comm.lib.enable_profiling()
    nprof = min(args.steps, 5)
    for _ in range(nprof):
        loop.step()
    phases = {k: v * 1e3 / nprof for k, v in loop.timer_collect().items()}
    prof = comm.lib.profile_report()
    comm.lib.disable_profiling()
    loop.timit = False
    kernels = {k: {"calls_per_step": n / nprof, "ms_per_step": t / nprof, "ms_per_call": t / n}
               for k, (n, t) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    particle_ms = sum(phases[k] for k in ("push-x", "sort", "depose", "gather + push-p"))

    # ---- roofline of the dominant kernel (largest share of the step)
    peaks, peak_src = measured_peaks()
    K = A["Nr"] - 1
    top = next(iter(kernels))
    roof = None
    # algorithmic bytes per particle of the reference stages each call stands for
    alg_bytes = {"chb_push_xyz": BYTES["push"], "chb_push_index": BYTES["push"] + 28,
                 "chb_index_and_sum": 28, "chb_sort_scatter_stable": 12,
                 "chb_depose_vector": BYTES["depose_vector"],
                 # fused: push_xyz + the whole first sort + depose_vector
                 "chb_push_depose_vector": BYTES["push"] + BYTES["sort"] + BYTES["depose_vector"],
                 # one pass: both half pushes, the first sort, depose_vector and the
                 # index/histogram pass (28 B) of the second sort
                 "chb_push_depose_push_index": 2 * BYTES["push"] + BYTES["sort"] +
                 BYTES["depose_vector"] + 28,
                 "chb_depose_scalar": BYTES["depose_scalar"], "chb_gather_push": BYTES["gather"]}
    rooflines = {}
    for name, kinfo in kernels.items():
        t = kinfo["ms_per_call"] * 1e-3
        if name in alg_bytes:
            ach = alg_bytes[name] * np_gpu / t / 1e9
            rooflines[name] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                               "peak_source": peak_src}
    dht_names = [k for k in kernels if k.startswith("chb_dht")]
    if dht_names:
        # FP64 contraction: denominator = cuBLAS DGEMM of the same shape, timed here
        a = torch.randn(K, K, dtype=torch.float64, device=dev)
        b = torch.randn(K, 2 * A["Nx"], dtype=torch.float64, device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        flop_c = 2.0 * K * K * 2 * A["Nx"]          # one complex right-hand side
        dgemm_tf = flop_c / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e12
        # FP64 tensor-pipe ceiling of this device, measured live (register-only DMMA loop)
        import ctypes
        scratch = torch.empty(148 * 512, dtype=torch.float64, device=dev)
        nfl = ctypes.c_double(0.0)
        st = torch.cuda.current_stream().cuda_stream
        from chimeracl_b200 import _lib as _clib
        _clib.check(comm.lib.chb_dmma_peak(scratch.data_ptr(), scratch.numel(), 2000,
                                           ctypes.byref(nfl), st), "chb_dmma_peak")
        e0.record()
        _clib.check(comm.lib.chb_dmma_peak(scratch.data_ptr(), scratch.numel(), 20000,
                                           ctypes.byref(nfl), st), "chb_dmma_peak")
        e1.record()
        torch.cuda.synchronize()
        dmma_tf = nfl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
        # per-step tally at M=1: 10 real + 19 complex contractions in the
        # reference; 3 of the complex ones are obtained here from a mirror identity
        Mm = A["M"]
        n_real, n_cplx = (10, 19) if Mm == 1 else ((10, 0) if Mm == 0 else (10, 35))
        # ... and the two contractions of m = 0 sources run on half of the kx columns
        n_cplx_exec = n_cplx - (3 if Mm >= 1 else 0) - (1 if Mm >= 1 and loop.real_m0_symmetry else 0)
        # with the kr-row sharded solve a rank executes 1/world of every contraction
        share = 1.0 / world if shard_solve else 1.0
        flops = (0.5 * n_real + n_cplx) * flop_c * share
        flops_exec = (0.5 * n_real + n_cplx_exec) * flop_c * share
        dht_ms = sum(kernels[k]["ms_per_step"] for k in dht_names)
        ach = flops / (dht_ms * 1e-3) / 1e12
        entry = {"bound": "tensor", "achieved": ach, "peak": dmma_tf, "unit": "TFLOP/s",
                 "frac": ach / dmma_tf, "traffic": None,
                 "flops_per_step": flops, "ms_per_step": dht_ms,
                 "executed_flops_per_step": flops_exec,
                 "executed_tflops": flops_exec / (dht_ms * 1e-3) / 1e12,
                 "cublas_dgemm_tflops": dgemm_tf,
                 "peak_source": "FP64 DMMA issue ceiling measured in this run "
                                "(chb_dmma_peak, register-only loop; MEASURED_PEAKS.json has no "
                                "FP64 entry); cuBLAS DGEMM %dx%dx%d in this run: %.1f TFLOP/s"
                                % (K, K, 2 * A["Nx"], dgemm_tf)}
        for k in dht_names:
            rooflines[k] = entry
        kernels["chb_dht*"] = {"calls_per_step": sum(kernels[k]["calls_per_step"] for k in dht_names),
                               "ms_per_step": dht_ms, "ms_per_call": dht_ms}
        rooflines["chb_dht*"] = entry
        top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    # DRAM bytes per launch from the committed ncu --set full capture of this build
    ncu_kernel = {"chb_dht*": "void dht_gemm_wide_kernel<7>", "chb_dht": "void dht_gemm_wide_kernel<7>",
                  "chb_dht2": "void dht_gemm_wide_kernel<7>",
                  "chb_dht_batched": "void dht_gemm_wide_kernel<7>",
                  "chb_gather_push": "void gather_push_kernel<1>",
                  "chb_push_depose_vector": "void depose_kernel<1, 1, 1>",
                  "chb_push_depose_push_index": "void depose_kernel<1, 1, 2, 32>",
                  "chb_depose_scalar": "void depose_kernel<1, 0, 0, 128>",
                  "chb_push_index": "void index_kernel<1>",
                  "chb_psatd_advance": "psatd_kernel",
                  "chb_fft_x_batched": "void fft_pow2_kernel<12>"}
    tpath = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if os.path.exists(tpath) and not args.small:
        with open(tpath) as f:
            tr = json.load(f)
        for name, entry in rooflines.items():
            v = tr.get(ncu_kernel.get(name, ""), {}).get("dram_bytes_per_launch")
            if shard_solve and name.startswith("chb_dht"):
                v = None            # captured for the unsharded launch shapes
            if v:
                entry["traffic"] = float(np.mean(v))
                entry["traffic_source"] = "profiles/r1_ncu_traffic.json (ncu --set full, per launch)"
                if entry["bound"] == "hbm" and name in kernels:
                    # measured DRAM traffic over the live launch time: the real HBM
                    # utilisation (the algorithmic figure above counts the bytes of the
                    # reference stages a fused call replaces and can exceed the peak)
                    gbs = entry["traffic"] / (kernels[name]["ms_per_call"] * 1e-3) / 1e9
                    entry["dram_gbs"] = gbs
                    entry["dram_frac"] = gbs / entry["peak"]
    roof = rooflines.get(top)
    if roof is None and rooflines:
        top = max(rooflines, key=lambda k: kernels[k]["ms_per_step"])
        roof = rooflines[top]
    if roof is not None:
        roof = dict(roof, kernel=top, share_of_step=kernels[top]["ms_per_step"] / ms_step)
    particle_roof = BYTES_PER_PARTICLE_STEP * np_gpu / (particle_ms * 1e-3) / 1e9

    # ---- end to end through the host-buffer API (rank-local, max over ranks)
    host_in, host_out = host_api.make_host_buffers(eons, solver)
    bytes_io = [0, 0]

    def e2e_step():
        bytes_io[0], bytes_io[1] = host_api.step_from_host(loop, eons, host_in, host_out)
    for _ in range(2):
        e2e_step()
    ne2e = max(3, min(args.steps, 5))
    ms_e2e = timed(e2e_step, ne2e)
    e2e_value = np_gpu * world * ne2e / (ms_e2e * 1e-3)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = run_reference(args, as_baseline=True)
        except Exception as exc:  # the baseline is a report, never a reason to fail
            cpu = {"error": repr(exc)}
    line = {"metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(bench_config(A, np_gpu, world),
                           **({"field_solve": "kr rows sharded x%d (all-gather rho/G spectra, "
                                              "all-reduce E/B partials)" % world}
                              if shard_solve else {})),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s",
                    "h2d_bytes_per_step": bytes_io[0], "d2h_bytes_per_step": bytes_io[1],
                    "ms_per_step": ms_e2e / ne2e},
            "roofline": roof, "cpu_baseline": cpu,
            "full_step_ms": ms_step, "particle_path_ms": particle_ms,
            "particle_path": {"value": np_gpu * world / (particle_ms * 1e-3),
                              "unit": "particle-steps/s (push+gather+sort+deposit only)",
                              "hbm_gbs_algorithmic": particle_roof,
                              "frac_of_hbm_peak": particle_roof / peaks["hbm_gbs"]},
            "phases_ms": phases, "kernels": kernels, "kernel_rooflines": rooflines}
    emit(line)


def emit(line):
    """Exactly one JSON line on the real stdout (libraries may print to fd 1)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)          # anything else written to stdout goes to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="debug-size grid (not a bench config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard-spectral", action="store_true",
                    help="(default for N>1) kr-row sharded field solve")
    ap.add_argument("--replicated-solve", action="store_true",
                    help="N>1: every rank runs the whole field solve (the round-1 baseline)")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        run_reference(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
