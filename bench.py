#!/usr/bin/env python
"""bench.py -- chimera-b200 headline benchmark.

Default workload (BASELINE.json configs[2], the configuration the metric is quoted on and
the largest that fits one GPU): uniform thermal plasma, Nx=4096, Nr=512, M=1 (modes 0,1),
16 particles per cell, electrons + immobile ions, DampCells=50, no laser, no frame.
A "step" is one full PIC_loop.step() (push+sort+deposit+transforms+PSATD+gather).
Metric: particle-steps/s = mobile particles (all ranks) x steps / time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--scaling strong|weak] [--config cfg3|cfg1|cfg2|cfg4|cfg5]

N > 1 (torchrun, one rank per GPU).  --scaling strong (the default, what configs[2] states:
"particles sharded over 1/2/4/8 B200"): the SAME 33.4 M electrons + ions are split over the
ranks (contiguous bands of the cell-sorted lattice); --scaling weak: every rank carries the
full 16 ppc shard.  Every rank holds the full grid, rho/J are summed with NCCL every step,
and the field solve is split over the ranks by kr rows (--replicated-solve: every rank
solves the whole grid).  For N > 1 the line carries `parity`: the sharded solve against the
replicated one on the same seeded state, and invariants of the summed deposit.
--impl reference times the reference's own CPU implementation (oracle/_ref = chimeraCL's
kernels host-compiled + OpenMP, np.dot, np.fft) on all host cores.
--config cfg1|cfg2|cfg4|cfg5: the other BASELINE configurations as extra lines (cfg2:
transformer round trip; cfg1/cfg4: the LWFA scripts; cfg5: the 10^9-particle shape).
"""
import os
import sys

# the host-compiled reference kernels (OpenMP) alternate with OpenBLAS calls: with libgomp's
# default active waiting the two thread pools fight for the cores (measured here: field
# phases 7x slower), so the CPU legs run with passive waiting -- the faster, fair setting
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
if "reference" in sys.argv:
    # the reference arm uses every host core it can, also under torchrun (which exports
    # OMP_NUM_THREADS=1): must be set before NumPy/OpenBLAS and libgomp are loaded
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import hashlib  # noqa: E402
import json  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# bytes per mobile particle-step, SURVEY.md 8(d): reads+writes each stage must move
BYTES = {"push": 80, "sort": 40, "depose_vector": 68, "depose_scalar": 36, "gather": 84}
BYTES_PER_PARTICLE_STEP = 2 * BYTES["push"] + 2 * BYTES["sort"] + BYTES["depose_vector"] \
    + BYTES["depose_scalar"] + BYTES["gather"]          # 428
FULL_NPPC = (2, 2, 4)


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


def species_cfgs(solver_args, nppc=FULL_NPPC):
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


def mobile_particles_total(A, nppc=FULL_NPPC):
    """In-domain particles of the lattice plasma_domain() creates: (Nx-3) x (Nr-2) cells,
    minus the radial positions of the top cell row that fall into grid row Nr-2 (particle
    cells are staggered half a cell from the grid rows: ir = floor(r/dr + 1/2))."""
    npx, npr, npt = nppc
    ncx, ncr = A["Nx"] - 3, A["Nr"] - 2
    lost_r = sum(1 for j in range(npr) if (j + 0.5) / npr >= 0.5)
    return ncx * npx * npt * (ncr * npr - lost_r)


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


def source_digest():
    """Digest of the CUDA sources the shipped library was built from: the committed ncu
    traffic table is only used when it was captured on the same sources."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "chimeracl_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def bench_config(A, np_total, n_gpus, scaling):
    per_gpu = np_total // n_gpus if scaling == "strong" else np_total
    cfg = {"workload": "uniform thermal plasma (BASELINE configs[2])", "Nx": int(A["Nx"]),
           "Nr": int(A["Nr"]), "modes": "m=0,1", "ppc": 16,
           "mobile_particles_total": int(per_gpu * n_gpus),
           "immobile_particles_total": int(per_gpu * n_gpus),
           "mobile_particles_per_gpu": int(per_gpu),
           "parallelism": "particles sharded x%d, grid replicated, NCCL all-reduce of rho/J"
                          % n_gpus,
           "start": "particles on the injector's lattice (16 per cell exactly), timed from step W "
                    "on, like the reference arm; later steps: see steady_state",
           "l2": "inputs exceed L2 (%.1f GB of particle data, 2.2 GB of fields per step)"
                 % (per_gpu * 96 / 1e9)}
    return cfg


# ============================================================================ reference arm
def run_reference(args, as_baseline=False):
    """chimeraCL's own CPU path on the host cores: its OpenCL C kernels compiled for
    the host (oracle/_ref, OpenMP over work-items) + np.dot + np.fft, driven by the
    restated methods/ orchestration.  Every step is a REAL full step on a bounded sample
    of the workload: the full grid (so the whole field solve), particles subsampled to
    2 per cell; `ms_per_step` is the measured wall time of such a step.  `value` is the
    throughput that implies for the full 16 ppc workload: the particle phases are linear
    in the particle count (x8), the field phases do not depend on it."""
    from oracle import orchestration as O
    from oracle.ref_kernels import RefKernels, ref_available
    from oracle.np_kernels import NumpyKernels
    cores = os.cpu_count() or 1
    O.FFT_WORKERS = cores            # x-FFTs over the rows on all cores (scipy.fft pocketfft)
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
    # (1, 2, 1): the same two radial positions per cell as the full (2, 2, 4) lattice, so
    # exactly 1/8 of its in-domain particles
    sample_nppc = (1, 2, 1)
    scale = int(np.prod(FULL_NPPC) // np.prod(sample_nppc))
    ecfg, icfg = species_cfgs(A, FULL_NPPC)
    dom = plasma_domain(A)
    rng = np.random.default_rng(1234)
    Nx_loc = int(np.ceil((dom["Xmax"] - dom["Xmin"]) / A["dx"]) + 1)
    Nr_loc = int(np.round((dom["Rmax"] - dom["Rmin"]) / A["dr"]) + 1)
    xg = dom["Xmin"] + A["dx"] * np.arange(Nx_loc)
    rg = dom["Rmin"] + A["dr"] * np.arange(Nr_loc)
    th = rng.uniform(0, 2 * np.pi, (Nx_loc - 1) * (Nr_loc - 1))
    x, y, z, w = NumpyKernels(M).fill_grid(th, xg, rg, sample_nppc)
    P = O.OracleParticles(ecfg, K)
    w = w * P.Args["w0"]
    px, py, pz = (rng.normal(0, 0.01, x.size) for _ in range(3))
    P.set_particles(x=x, y=y, z=z, px=px, py=py, pz=pz, w=w,
                    g_inv=1 / np.sqrt(1 + px * px + py * py + pz * pz))
    I = O.OracleParticles(icfg, K)
    I.set_particles(x=x.copy(), y=y.copy(), z=z.copy(), w=w.copy())
    species = [P, I]
    for p in species:            # what the GPU arm does: drop the out-of-domain lattice rows
        p.sort_parts(S)
        p.align_parts()
    n = int(P.Args["Np"])
    np_full = n * scale
    assert args.small or np_full == mobile_particles_total(A), (np_full, mobile_particles_total(A))

    def one_step():
        t0 = time.perf_counter()
        for p in species:
            p.push_coords("half")
            p.sort_parts(S)
        S.depose_currents(species)
        for p in species:
            p.push_coords("half")
            p.sort_parts(S)
        S.depose_charge(species)
        t1 = time.perf_counter()
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
        t2 = time.perf_counter()
        S.gather_and_push(species)
        t3 = time.perf_counter()
        return (t1 - t0) + (t3 - t2), (t2 - t1)

    steps = max(1, args.steps if not as_baseline else 3)
    warm = args.warmup if not as_baseline else 1
    for i in range(warm):
        one_step()
    tp = tf = 0.0
    t_wall = time.perf_counter()
    for i in range(steps):
        a, b = one_step()
        tp += a
        tf += b
    t_wall = time.perf_counter() - t_wall
    t_step_full = (tp * scale + tf) / steps          # what a 16 ppc step costs
    value = np_full / t_step_full
    sample = ("full grid Nx=%d Nr=%d M=1, %d real steps; particles subsampled to %d per cell "
              "(%d mobile + as many ions = 1/%d of the workload); measured per step: particle "
              "phases %.3f s, field phases %.3f s; value = %d / (%d x particle + field)" %
              (A["Nx"], A["Nr"], steps, int(np.prod(sample_nppc)), n, scale,
               tp / steps, tf / steps, np_full, scale))
    cpu = {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind,
           "sample": sample, "threads_env": os.environ.get("OMP_NUM_THREADS")}
    if as_baseline:
        return cpu
    scaling = args.scaling or ("strong" if args.gpus > 1 else "weak")
    line = {"impl": "reference", "metric": "particle-steps/s", "value": value,
            "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": t_wall / steps * 1e3,
            "ms_per_step_full_workload": t_step_full * 1e3,
            "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(A, np_full, args.gpus, scaling),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    emit(line)


# ============================================================================ our arm
def build_case(comm, grid, scaling, seed, shard_solve):
    """Solver + electrons + ions of the cfg3 workload for this rank."""
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    solver = Solver(dict(grid), comm)
    A = solver.Args
    ecfg, icfg = species_cfgs(A)
    eons, ions = Particles(ecfg, comm), Particles(icfg, comm)
    ions.Args["InjectorSource"] = eons
    dom = plasma_domain(A)
    if scaling == "strong" and comm.world_size > 1:
        # this rank's contiguous band of the cell-sorted lattice (equal counts per rank)
        dom["r_shard"] = (comm.rank, comm.world_size)
    comm.generator.manual_seed(seed)
    eons.make_new_domain(dom)
    eons.add_new_particles()
    ions.add_new_particles(source=eons)
    eons.free_added()
    for p in (eons, ions):
        p.sort_parts(solver)
        p.align_parts()
    if shard_solve:
        solver.enable_spectral_sharding()
    return solver, eons, ions


def deposit_invariants(solver, eons, world):
    """Checks of the summed charge deposit that hold for any number of ranks: the number
    of in-domain electrons, and the grid charge of the electrons alone against the sum of
    their weights (corrected for the ghost-row rule of the reference: the share deposited
    on row 0 is SUBTRACTED from row 1, kernels/grid_generic.cl:44)."""
    import torch
    import torch.distributed as dist
    A = solver.Args
    solver.depose_charge(species=[eons])
    rho = solver.DataDev["rho_m0"].t
    dv = torch.from_numpy(A["dV_inv"]).to(rho.device)
    q_grid = float((rho[1:] / dv[1:, None]).sum().item())
    x, y, z, w = (eons.DataDev[k].t for k in ("x", "y", "z", "w"))
    r = torch.sqrt(y * y + z * z)
    ix = torch.floor((x - A["Xmin"]) * A["dx_inv"])
    ir = torch.floor((r - A["Rmin"]) * A["dr_inv"])
    ok = (ix > 0) & (ix < A["Nx"] - 2) & (ir < A["Nr"] - 2)
    ghost = torch.where(ok & (ir == 0), w * (0.5 - r * A["dr_inv"]), torch.zeros_like(w))
    t = torch.stack(((w * ok).sum(), ghost.sum(), ok.sum().double()))
    if world > 1:
        dist.all_reduce(t)
    q_part = -(float(t[0]) - 2 * float(t[1]))
    return {"np_stay_total": int(t[2].item()),
            "np_expected": mobile_particles_total(A) if A["Nx"] == 4096 else None,
            "charge_rel_err": abs(q_grid - q_part) / abs(q_part)}


def sharded_vs_replicated(comm, grid, scaling, seed, steps=2):
    """N > 1: `steps` PIC steps from the same seeded state with the kr-row sharded solve
    and with the replicated one; largest difference over all ranks of the E / B grids
    (relative to each field's maximum) and of the electron momenta (per particle)."""
    import torch
    import torch.distributed as dist
    from chimeracl_b200.pic_loop import PIC_loop
    res = []
    for sharded in (False, True):
        solver, eons, ions = build_case(comm, grid, scaling, seed, sharded)
        loop = PIC_loop(solvers=[solver], species=[eons, ions])
        for _ in range(steps):
            loop.step()
        comm.synchronize()
        out = {k: solver.DataDev[k].t.clone() for k in solver.DataDev
               if k[0] in "EB" and k[1] in "xyz" and "_fb_" not in k}
        for k in ("px", "py", "pz"):
            out[k] = eons.DataDev[k].t.clone()
        res.append(out)
        del solver, eons, ions, loop
    ref, got = res
    worst = {"E": 0.0, "B": 0.0}
    for f in "EB":
        scale = max(float(ref[k][1:].abs().max()) for k in ref if k[0] == f)
        for k in ref:
            if k[0] == f:
                worst[f] = max(worst[f], float((got[k][1:] - ref[k][1:]).abs().max()) / scale)
    d2 = sum((got[k] - ref[k]) ** 2 for k in ("px", "py", "pz"))
    n2 = sum(ref[k] ** 2 for k in ("px", "py", "pz"))
    wp = float(torch.sqrt(d2 / n2.clamp_min(1e-300)).max()) if d2.numel() else 0.0
    t = torch.tensor([worst["E"], worst["B"], wp], dtype=torch.float64, device=comm.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"max_rel_E": float(t[0]), "max_rel_B": float(t[1]), "max_rel_p": float(t[2]),
            "steps": steps, "against": "replicated field solve, same seeded state, NCCL"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from chimeracl_b200 import _lib
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    from chimeracl_b200.pic_loop import PIC_loop, loop_steps
    from chimeracl_b200.parallel import init_distributed
    from chimeracl_b200 import host_api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm = Communicator(answers=[0, 0], seed=1234 + rank)
    init_distributed(comm)
    dev = comm.device
    scaling = args.scaling or ("strong" if world > 1 else "weak")

    grid = workload(args.small)
    # N > 1: kr-row sharded field solve; --replicated-solve or
    # CHB_SHARD_SPECTRAL=0 brings the replicated solve back
    shard_solve = world > 1 and not args.replicated_solve and \
        os.environ.get("CHB_SHARD_SPECTRAL", "1") != "0"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if world > 1 and shard_solve and not args.no_parity:
        parity = sharded_vs_replicated(comm, grid, scaling, 4321 + rank)
        torch.cuda.empty_cache()

    solver, eons, ions = build_case(comm, grid, scaling, 1234 + rank, shard_solve)
    A = solver.Args
    np_gpu = int(eons.Args["Np"])
    cnt = torch.tensor([np_gpu], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    np_all = int(cnt.item())
    loop = PIC_loop(solvers=[solver], species=[eons, ions])

    def timed(fn, steps, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
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
    value = np_all * args.steps / (ms * 1e-3)

    # ---- phase split and per-kernel times (CUDA events around every C-ABI call)
    loop.timit = True
    loop.Timer = {k: 0 for k in loop_steps}
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
    # compulsory bytes per particle of the fused passes themselves: 8 attributes + sort_indx
    # read, x y z + cell index written
    own_bytes = {"chb_push_depose_push_index": 64 + 4 + 24 + 4, "chb_push_depose_vector": 64 + 4 + 24}
    rooflines = {}
    for name, kinfo in kernels.items():
        t = kinfo["ms_per_call"] * 1e-3
        if name in alg_bytes:
            ach = alg_bytes[name] * np_gpu / t / 1e9
            rooflines[name] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                               "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                               "peak_source": peak_src}
            if name in own_bytes:
                # a fused call is credited above with the bytes of every reference stage it
                # replaces (can exceed the peak); this is what the fused pass itself must move
                own = own_bytes[name] * np_gpu / t / 1e9
                rooflines[name].update(fused_pass_bytes_per_particle=own_bytes[name],
                                       fused_pass_gbs=own, fused_pass_frac=own / peaks["hbm_gbs"])
    dht_names = [k for k in kernels if k.startswith("chb_dht")]
    if dht_names:
        # FP64 contraction: cuBLAS DGEMM of the same shape, timed here, for comparison
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
        _lib.check(comm.lib.chb_dmma_peak(scratch.data_ptr(), scratch.numel(), 2000,
                                          ctypes.byref(nfl), st), "chb_dmma_peak")
        e0.record()
        _lib.check(comm.lib.chb_dmma_peak(scratch.data_ptr(), scratch.numel(), 20000,
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
        # `achieved` / `frac`: the flops the kernels really execute over the measured
        # ceiling of the FP64 tensor pipe; the reference's own tally (4 contractions more,
        # saved here by identities) is reported beside it
        ach = flops_exec / (dht_ms * 1e-3) / 1e12
        entry = {"bound": "tensor", "achieved": ach, "peak": dmma_tf, "unit": "TFLOP/s",
                 "frac": ach / dmma_tf, "traffic": None,
                 "executed_flops_per_step": flops_exec, "ms_per_step": dht_ms,
                 "reference_tally_flops_per_step": flops,
                 "reference_tally_tflops": flops / (dht_ms * 1e-3) / 1e12,
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
    # DRAM bytes per launch from the committed ncu --set full capture -- only if it was
    # taken on the sources this library was built from (tools/ncu_traffic.py records the
    # digest), so that the figure cannot silently go stale
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic_note = "no ncu capture committed"
    if os.path.exists(tpath) and not args.small:
        with open(tpath) as f:
            tr = json.load(f)
        if tr.get("source_digest") != source_digest():
            traffic_note = ("profiles/ncu_traffic.json was captured on other kernel sources "
                            "(digest %s, now %s): not used" % (tr.get("source_digest"), source_digest()))
        else:
            traffic_note = "profiles/ncu_traffic.json (ncu --set full, per launch, digest matches)"
            for name, entry in rooflines.items():
                v = tr.get("calls", {}).get("chb_dht_batched" if name == "chb_dht*" else name)
                if shard_solve and name.startswith("chb_dht"):
                    v = None            # captured for the unsharded launch shapes
                if scaling == "strong" and world > 1 and entry["bound"] == "hbm":
                    v = None            # captured with the whole 33.4 M particles on one GPU
                if v:
                    entry["traffic"] = float(v["dram_bytes_per_launch"])
                    entry["traffic_kernel"] = v.get("kernel")
                    if entry["bound"] == "hbm" and name in kernels:
                        # measured DRAM traffic over the live launch time: the real HBM
                        # utilisation (the algorithmic figure counts the bytes of the
                        # reference stages a fused call replaces and can exceed the peak)
                        gbs = entry["traffic"] / (kernels[name]["ms_per_call"] * 1e-3) / 1e9
                        entry["dram_gbs"] = gbs
                        entry["dram_frac"] = gbs / entry["peak"]
    roof = rooflines.get(top)
    if roof is None and rooflines:
        top = max(rooflines, key=lambda k: kernels[k]["ms_per_step"])
        roof = rooflines[top]
    particle_roof = BYTES_PER_PARTICLE_STEP * np_gpu / (particle_ms * 1e-3) / 1e9
    if roof is not None:
        roof = dict(roof, kernel=top, share_of_step=kernels[top]["ms_per_step"] / ms_step,
                    traffic_source=traffic_note,
                    particle_path={"ms_per_step": particle_ms,
                                   "value": np_all / (particle_ms * 1e-3),
                                   "unit": "particle-steps/s (push+gather+sort+deposit only)",
                                   "hbm_gbs_algorithmic": particle_roof,
                                   "frac_of_hbm_peak": particle_roof / peaks["hbm_gbs"],
                                   "bytes_per_particle_step": BYTES_PER_PARTICLE_STEP})

    invariants = deposit_invariants(solver, eons, world)

    # ---- end to end through the host-buffer API: every step uploads all particle
    # attributes from pinned host memory and downloads the updated ones + rho_m0; the
    # streaming pipeline overlaps step k+1's upload with step k's compute and download
    host_in, host_out = host_api.make_host_buffers(eons, solver)
    host_out2 = {k: torch.empty_like(v).pin_memory() for k, v in host_out.items()}
    bytes_io = [0, 0]
    pipe = host_api.HostStepPipeline(loop, eons, depth=int(os.environ.get("CHB_E2E_DEPTH", "2")))
    outs = (host_out, host_out2)

    def e2e_step():
        bytes_io[0], bytes_io[1] = pipe.submit(host_in, outs[pipe.k % 2])
    for _ in range(3):
        e2e_step()
    ne2e = max(5, args.steps)
    ms_e2e = timed(e2e_step, ne2e, after=pipe.drain)
    e2e_value = np_all * ne2e / (ms_e2e * 1e-3)
    # and one step at a time (upload -> step -> download, nothing overlapped across steps)
    ms_e2e_single = timed(lambda: host_api.step_from_host(loop, eons, host_in, host_out), 3) / 3

    # ---- drift of the storage order (reference semantics: particles stay in storage order
    # and are visited through sort_indx; only align_parts() re-sorts the storage, and the
    # reference calls it on plasma injection only, frame.py:59).  The headline above is
    # measured like the reference arm, from the lattice start; here the same loop further
    # into the run, and with PIC_loop(align_every=10) calling align_parts() periodically
    steady = None
    if not args.no_steady_state:
        steady = {"note": "ms per PIC step later in the same run; align_every: PIC_loop option "
                          "(opt-in) that calls the reference's align_parts() every N steps"}
        for _ in range(100):
            loop.step()
        steady["no_align_ms_per_step"] = timed(loop.step, 10) / 10
        steady["no_align_steps_since_start"] = int(loop.it)
        loop.align_every = 10
        for _ in range(20):
            loop.step()
        steady["align_every_10_ms_per_step"] = timed(loop.step, 20) / 20
        loop.align_every = 0

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
    if parity is None:
        parity = {}
    parity.update(invariants)
    line = {"metric": "particle-steps/s", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(A, np_all if scaling == "strong" else np_gpu, world, scaling),
            "field_solve": ("kr rows sharded x%d (all-gather rho/G spectra, all-reduce E/B "
                            "partials)" % world) if shard_solve else "replicated",
            "exchange": (("own kernels over NVLink peer memory (chb_peer_*: %s), exchange stream"
                          % ("NVSwitch multimem" if any(v[2] for v in solver._sharding["peer"].values())
                             else "P2P loads/stores"))
                         if shard_solve and "peer" in solver._sharding else
                         ("NCCL" if world > 1 else None)),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s",
                    "h2d_bytes_per_step": bytes_io[0], "d2h_bytes_per_step": bytes_io[1],
                    "ms_per_step": ms_e2e / ne2e, "steps": ne2e,
                    "mode": "streaming (host_api.HostStepPipeline: upload of step k+1 under "
                            "compute and download of step k)",
                    "ms_per_step_unpipelined": ms_e2e_single},
            "roofline": roof, "cpu_baseline": cpu, "parity": parity, "steady_state": steady,
            "full_step_ms": ms_step, "particle_path_ms": particle_ms,
            "phases_ms": phases, "kernels": kernels, "kernel_rooflines": rooflines}
    emit(line)


# ============================================================================ other configs
def run_other_config(args):
    """BASELINE configs other than the one the metric is quoted on, one JSON line each
    (extra evidence; the headline line is cfg3's)."""
    import torch
    from chimeracl_b200.methods.generic_methods_cl import Communicator
    from chimeracl_b200.parallel import init_distributed
    from chimeracl_b200.particles import Particles
    from chimeracl_b200.solver import Solver
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    comm = Communicator(answers=[0, 0], seed=1234 + rank)
    init_distributed(comm)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    if args.config == "cfg2":
        # examples/test_transformer.py:9-49 at Nx=2048, Nr=256, M=1
        grid_in = {"Xmin": -1., "Xmax": 1., "Nx": 2048, "Rmin": 0, "Rmax": 1., "Nr": 256, "M": 1}
        parts = Particles(dict(grid_in), comm)
        grid = Solver(dict(grid_in), comm)
        parts.add_particles(beam_in={"Np": int(7e6), "FullCharge": 1, "x_c": 0., "Lx": 0.3,
                                     "y_c": 0.2, "Ly": 0.3, "z_c": 0.2, "Lz": 0.3})
        parts.sort_parts(grid=grid)
        parts.align_parts()
        grid.depose_charge([parts])
        tmp0 = grid.DataDev["rho_m0"].get().copy()
        tmp1 = grid.DataDev["rho_m1"].get().copy()
        grid.fb_transform(scals=["rho"], dir=0)
        grid.set_to(grid.DataDev["rho_m0"], 0)
        grid.set_to(grid.DataDev["rho_m1"], 0)
        grid.fb_transform(scals=["rho"], dir=1)
        err = float((np.abs(grid.DataDev["rho_m0"].get() - tmp0)[1:] / np.abs(tmp0[1:]).max()
                     + np.abs(grid.DataDev["rho_m1"].get() - tmp1)[1:] / np.abs(tmp1[1:]).max()).max())
        for _ in range(max(args.warmup, 3)):
            grid.fb_transform(scals=["rho"], dir=0)
            grid.fb_transform(scals=["rho"], dir=1)
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            grid.fb_transform(scals=["rho"], dir=0)
            grid.fb_transform(scals=["rho"], dir=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        emit({"metric": "ms per forward+backward Fourier-Bessel transform pair (rho, m=0,1)",
              "value": ms, "unit": "ms", "n_gpus": 1, "steps": args.steps,
              "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": False,
              "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": "transformer round trip (BASELINE configs[1], "
                                     "examples/test_transformer.py)", "Nx": 2048, "Nr": 256,
                         "modes": "m=0,1", "beam_particles": int(7e6)},
              "round_trip_error": err, "round_trip_tolerance": 1e-12})
        return

    # cfg1 / cfg4 / cfg5: the LWFA scripts
    import importlib.util
    name = "lpa_script_small" if args.config == "cfg1" else "lpa_script_large"
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if args.config == "cfg1":
        _, solver, eons, ions, frame, loop = mod.build(comm=comm)
        what = "examples/lpa_script_small.py (BASELINE configs[0]): Nx=900 Nr=90 M=1"
    else:
        solver, eons, ions, loop = mod.build(comm, cfg5=(args.config == "cfg5"),
                                             replicated_solve=args.replicated_solve)
        what = ("examples/lpa_script_large.py (BASELINE configs[3]): Nx=4096 Nr=252 M=1"
                if args.config == "cfg4" else
                "BASELINE configs[4]: Nx=16384 Nr=1024 M=2, 32 ppc, plasma pre-filled")
    import torch.distributed as dist
    nwarm = max(args.warmup, 3) if args.config == "cfg5" else 200   # LWFA: let plasma enter
    for _ in range(nwarm):
        loop.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        loop.step()
    e1.record()
    torch.cuda.synchronize()
    # the same steps replayed from CUDA graphs (PIC_loop(use_cuda_graph=True); single rank):
    # these configurations are launch-bound, the timed steps include the eager steps and
    # the re-capture every plasma injection forces
    graph = None
    if world == 1:
        loop.use_cuda_graph = True
        for _ in range(4):
            loop.step()
        torch.cuda.synchronize()
        r0 = loop.graph_replays
        g0, g1 = ev(), ev()
        g0.record()
        for _ in range(args.steps):
            loop.step()
        g1.record()
        torch.cuda.synchronize()
        graph = {"ms_per_step": g0.elapsed_time(g1) / args.steps,
                 "replayed_steps": loop.graph_replays - r0, "of_steps": args.steps,
                 "captures_total": loop.graph_captures}
        loop.use_cuda_graph = False
    t = torch.tensor([e0.elapsed_time(e1) / args.steps, float(eons.Args["Np"])],
                     dtype=torch.float64, device=comm.device)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t[0] = tm[0]
    ms, n_all = float(t[0]), int(t[1])
    finite = all(bool(torch.isfinite(solver.DataDev[k].t.abs().sum()).item())
                 for k in ("Ez_m0", "rho_m0"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit({"metric": "particle-steps/s", "value": n_all / (ms * 1e-3),
              "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": nwarm,
              "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": {"workload": what, "mobile_particles_total": n_all,
                         "steps_before_timing": nwarm}, "fields_finite": finite,
              "cuda_graph": graph})


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
    ap.add_argument("--config", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="N>1: split cfg3's particles over the ranks (strong, default) or "
                         "give every rank the full 16 ppc shard (weak)")
    ap.add_argument("--small", action="store_true", help="debug-size grid (not a bench config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-steady-state", action="store_true",
                    help="skip the extra 140 steps that measure the storage-order drift")
    ap.add_argument("--no-parity", action="store_true",
                    help="N>1: skip the sharded-vs-replicated parity steps")
    ap.add_argument("--replicated-solve", action="store_true",
                    help="N>1: every rank runs the whole field solve (the round-1 baseline)")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        run_reference(args)
        return
    if args.config != "cfg3":
        run_other_config(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
