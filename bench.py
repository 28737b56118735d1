#!/usr/bin/env python
"""bench.py -- headline benchmark of the NDSM vector-potential hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config dipole513|charges1025x1025x257|poisson_weak] [--size n]

A "step" is one complete solve of the configuration's workload, FP64, default tolerances:
  dipole513 (default, the configuration BASELINE.json's metric is quoted on): vector-potential solve (BC setup:
      6 chi solves -> 3 multigrid solves to vc_tol -> flux-balance fields -> curl) of the synthetic sub-surface
      dipole on 513^3 (--size n for n^3); strong scaling over the GPUs;
  charges1025x1025x257 (BASELINE config 4): the same solve of a bipolar two-charge magnetogram on the
      non-cubic 1025 x 1025 x 257 box, mean metric; strong scaling;
  poisson_weak (BASELINE config 5): the scalar Poisson backend, 1025 x 1025 x (128 G + 1) on G GPUs, copt
      NDDNDD, analytic right-hand side; weak scaling.

value   : fine-grid Gpoint-updates/s of the device-resident solve (inputs already in HBM), i.e.
          sum_c(V-cycles_c * 2*ms * nx*ny*nz) / time  (SURVEY.md 8d), whole job over all N GPUs.
e2e     : the same metric through the reference-facing C ABI with HOST buffers (pinned), H2D of the inputs and
          D2H of the results inside the timed region (`ndsm_vector_solve` at N = 1).
roofline: the dominant kernel (k_relax3d colour pass on the finest level) timed live with CUDA events on the
          library's stream; algorithmic bytes = 8 B (rhs == 0) or 12 B per fine-grid point per launch.
cpu_baseline / --impl reference: the CPU oracle (restatement of the reference's OpenMP path; the Fortran
          reference cannot be compiled in this image) on the SAME configuration with all physical host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
METRIC = "fine-grid Gpoint-updates/s (time-to-vc_tol in ms_per_step)"


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "250"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        for ts, r in self.rows:
            if ts < t0 or ts > t1 + 0.3:
                continue  # the sampler runs from before the warm-up; only samples of the timed region count
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# configurations (the `config` object is identical in both arms)
# ------------------------------------------------------------------------------------------------
class Config:
    def __init__(self, args, world):
        self.name = args.config
        self.world = world
        if self.name == "dipole513":
            n = args.n
            self.shape = (n, n, n)
            self.mean = False
            self.scaling = "strong"
            self.workload = ("dipole-%d^3 vector_potential, default options (max metric, vc_tol=1e-10, ex_tol=1e-13, "
                             "ms=5)" % n)
        elif self.name == "charges1025x1025x257":
            self.shape = (1025, 1025, 257)
            self.mean = True
            self.scaling = "strong"
            self.workload = ("two-charge magnetogram 1025x1025x257 vector_potential (BASELINE config 4), mean metric, "
                             "vc_tol=1e-10, ex_tol=1e-13, ms=5")
        else:
            self.shape = (1025, 1025, 128 * world + 1)
            self.mean = False
            self.scaling = "weak"
            self.workload = ("scalar Poisson backend (ndsm_poisson) 1025x1025x(128*G+1), copt NDDNDD, analytic rhs, max "
                             "metric, vc_tol=1e-10, ex_tol=1e-13, ms=5 (BASELINE config 5)")
        self.npoints = self.shape[0] * self.shape[1] * self.shape[2]

    def describe(self):
        nx, ny, nz = self.shape
        return {"workload": self.workload, "shape": [nx, ny, nz],
                "l2": ("inputs larger than L2 (each %dx%dx%d fp64 array = %.2f GB)" if 8 * self.npoints > 126e6 else
                       "arrays of %dx%dx%d fp64 = %.3f GB fit the 126 MB L2: not a bandwidth measurement")
                      % (nx, ny, nz, 8 * self.npoints / 1e9)}

    def mesh(self):
        from ndsm_b200 import synthetic
        return synthetic.mesh(*self.shape)

    def field(self, x, y, z, faces_only=True):
        from ndsm_b200 import synthetic
        if self.name == "charges1025x1025x257":
            return synthetic.charges(x, y, z, faces_only=faces_only)
        return synthetic.dipole(x, y, z, faces_only=faces_only)

    def poisson_problem(self, k0, k1):
        """planes [k0,k1) of u_exact and rhs of config 5: u = cos(pi x) sin(pi y/Ly) sin(pi z/Lz)."""
        x, y, z = self.mesh()
        Ly, Lz = y[-1], z[-1]
        cx, sy, sz = np.cos(np.pi * x), np.sin(np.pi * y / Ly), np.sin(np.pi * z[k0:k1] / Lz)
        uex = sz[:, None, None] * sy[None, :, None] * cx[None, None, :]
        lam = -(np.pi ** 2) * (1.0 + 1.0 / Ly ** 2 + 1.0 / Lz ** 2)
        return uex, lam * uex


def updates_from_cycles(cyc, n_points, ms=5):
    """sum over Ax,Ay,Az of V-cycles * 2*ms * N.  Az always uses ms=5 (reference quirk)."""
    return sum(cyc[6 + c] * 2 * (5 if c == 2 else ms) * n_points for c in range(3))


def trace_cycles(lib, dist=None):
    """V-cycle counts of the last solve (chi1..6, Ax, Ay, Az).  Multi-GPU: a chi solve is only traced on the rank
    that owns the face, so the counts are combined with a max over the ranks."""
    cyc = [lib.ndsm_b200_trace_ncycles(s) for s in range(9)]
    if dist is not None:
        import torch
        t = torch.tensor(cyc, dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cyc = [int(v) for v in t.tolist()]
    return cyc


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle): the reference's algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def physical_cores():
    """Physical cores this process may use (OMP_NUM_THREADS = physical cores)."""
    try:
        allowed = os.sched_getaffinity(0)
    except Exception:
        allowed = set(range(os.cpu_count() or 1))
    cores = set()
    try:
        cpu, phys, core = None, None, None
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("processor"):
                    cpu = int(line.split(":")[1])
                elif line.startswith("physical id"):
                    phys = int(line.split(":")[1])
                elif line.startswith("core id"):
                    core = int(line.split(":")[1])
                elif not line.strip():
                    if cpu is not None and cpu in allowed and phys is not None and core is not None:
                        cores.add((phys, core))
                    cpu, phys, core = None, None, None
    except Exception:
        cores = set()
    n = len(cores) if cores else len(allowed)
    return max(1, n)


def cpu_env():
    """Pin the OpenMP run-time of the CPU arm (before libgomp is loaded): the launcher's OMP_NUM_THREADS=1 of
    torch.distributed.run must not leak into the baseline."""
    n = physical_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["OMP_PROC_BIND"] = "close"
    os.environ["OMP_PLACES"] = "cores"
    return n


def cpu_oracle(threads):
    from oracle import pyoracle as O
    O.set_num_threads(threads)   # also when libgomp was initialised earlier in this process
    return O


def cpu_vcycle_sample(cfg, ncycles, threads):
    """`ncycles` V-cycles (+ update_u) of the configuration's first 3D solve on the full mesh with the oracle.
    Returns (Gpoint-updates/s, seconds, threads, description)."""
    O = cpu_oracle(threads)
    x, y, z = cfg.mesh()
    nx, ny, nz = cfg.shape
    if cfg.name == "poisson_weak":
        copt = "NDDNDD"
        uex, rhs = cfg.poisson_problem(0, nz)
        u0 = np.zeros_like(rhs)
        del uex
    else:
        copt = "NDDNDD"   # the Ax solve: Dirichlet data on the y and z faces, rhs = 0
        u0 = np.zeros((nz, ny, nx))
        yy, xx = np.meshgrid(y, x, indexing="ij")
        u0[0, :, :] = np.sin(np.pi * xx) * np.cos(np.pi * yy)
        u0[-1, :, :] = 0.1 * u0[0, :, :]
        u0[:, 0, :] = u0[0, 0, :][None, :]
        u0[:, -1, :] = u0[0, -1, :][None, :]
        rhs = np.zeros_like(u0)
    mg = O.OracleMG([x, y, z], copt, ms=5)
    mg.load(u0, rhs)
    t0 = time.perf_counter()
    for _ in range(ncycles):       # solve_poisson_bvp's loop body: V-cycle + update_u
        mg.v_cycle()
        O.update_u(mg.u(0), u0)
    dt = time.perf_counter() - t0
    mg.close()
    upd = ncycles * 2 * 5 * cfg.npoints
    return upd / dt / 1e9, dt, O.num_threads(), ("%d V-cycle(s) + update_u of the first 3D solve (copt %s, ms=5) on the "
                                                  "full %dx%dx%d mesh" % (ncycles, copt, nx, ny, nz))


def cpu_full_solve(cfg, threads):
    """One complete solve of the configuration through the oracle's copy of the reference ABI.
    Returns (Gpoint-updates/s, seconds, threads, description, cycles)."""
    O = cpu_oracle(threads)
    x, y, z = cfg.mesh()
    if cfg.name == "poisson_weak":
        nz = cfg.shape[2]
        uex, rhs = cfg.poisson_problem(0, nz)
        t0 = time.perf_counter()
        ierr, u, du, nc = O.poisson_solve([x, y, z], "NDDNDD", np.zeros_like(rhs), rhs)
        dt = time.perf_counter() - t0
        upd = nc * 2 * 5 * cfg.npoints
        return upd / dt / 1e9, dt, O.num_threads(), "1 full solve_poisson_bvp (%d V-cycles)" % nc, [nc]
    b = cfg.field(x, y, z)
    t0 = time.perf_counter()
    ierr, A, B, tr = O.vector_potential(x, y, z, b, mean=cfg.mean, trace=True)
    dt = time.perf_counter() - t0
    names = ["chi%d" % f for f in range(1, 7)] + ["Ax", "Ay", "Az"]
    cyc = [len(tr[k]["du"]) for k in names]
    upd = updates_from_cycles(cyc, cfg.npoints)
    return (upd / dt / 1e9, dt, O.num_threads(),
            "1 full vector_potential solve through the oracle's ndsm_vector_solve ABI (V-cycles Ax/Ay/Az %d/%d/%d)"
            % (cyc[6], cyc[7], cyc[8]), cyc)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores.  The Fortran
    reference cannot be built in this image (no gfortran), so this is the oracle port (kind "port") with all
    physical cores, on the SAME mesh and configuration as the GPU arm at every N.  One step = one complete solve
    when K+W of them fit the time budget (small meshes); otherwise ONE complete solve is timed (it is ~80 s at
    513^3, so K and W are clamped to 1 and 0 and the line says so) -- the metric is a rate and the time-to-vc_tol
    of one solve, neither depends on K."""
    if rank != 0:
        return
    threads = cpu_env()
    cfg = Config(args, world)
    budget_s = float(os.environ.get("NDSM_BENCH_CPU_BUDGET_S", "240"))
    # probe: one V-cycle tells how long a solve takes (~45 cycles + BC setup)
    v1, dt1, thr, _ = cpu_vcycle_sample(cfg, 1, threads)
    est_full = dt1 * (16 if cfg.name == "poisson_weak" else 46)
    if est_full > 4 * budget_s:
        # far beyond the budget (1025-class meshes on few cores): a bounded number of V-cycles of the same mesh
        ncyc = max(1, int(budget_s / 2 / dt1))
        value, dt, thr, desc = cpu_vcycle_sample(cfg, ncyc, threads)
        steps_done, clamp = 1, "bounded sample: %s (a full solve would take ~%.0f s)" % (desc, est_full)
        ms_per_step = dt * 1e3
    else:
        nrep = args.warmup + args.steps
        timed = args.steps
        if est_full * nrep > budget_s:
            nrep, timed = 1, 1
        vals, times = [], []
        for i in range(nrep):
            v, dt, thr, desc, cyc = cpu_full_solve(cfg, threads)
            if i >= nrep - timed:
                vals.append(v * dt)
                times.append(dt)
        value = float(np.sum(vals) / np.sum(times))
        ms_per_step = float(np.mean(times) * 1e3)
        steps_done = timed
        clamp = desc if timed == args.steps else desc + "; K,W clamped to 1,0 (one CPU solve takes %.0f s)" % times[0]
    cpu = {"value": value, "unit": "Gpoint-updates/s", "cores": thr, "kind": "port", "sample": clamp,
           "omp": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OMP_PROC_BIND", "OMP_PLACES")}}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gpoint-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_timed": steps_done, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": cfg.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg.describe(), "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "Gpoint-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
PROF_NAMES = ["k_relax3d colour pass (finest level)", "k_residual3d", "restriction (finest -> level 1, k_restrict_direct)",
              "k_interp_add_zt", "update_u (k_diff_partial+final)", "halo exchange", "levels >= 2 of one V-cycle",
              "level 1 of one V-cycle (2 brackets per cycle)"]


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the finest-level colour pass, from the committed
    `ncu --set full` capture under profiles/ (keyed by points per launch); None when no capture matches."""
    out = {}
    try:
        with open(os.path.join(REPO, "profiles", "ncu_traffic.json")) as f:
            for k, v in json.load(f).get("k_relax3d", {}).items():
                out[int(k)] = float(v)
    except Exception:
        pass
    return out


def kernel_report(lib, npoints, dev_ms, rhs_zero=True, single_gpu=True):
    """Per-kernel-class CUDA-event timings gathered by the library (finest level of this rank's slab)."""
    cnt, tot = ctypes.c_ulonglong(0), ctypes.c_double(0)
    peak, peak_src = measured_peak()
    kern = {}
    b0 = 8.0 if rhs_zero else 12.0
    # SURVEY 8d / DESIGN.md (level 0); update_u on one GPU only measures the difference (ping-pong cycles, no copy): 16 B/pt
    bytes_per_pt = [b0, 16.0 if rhs_zero else 24.0, 9.0, 17.0, 16.0 if single_gpu else 24.0, 0.0, 0.0, 0.0]
    for cls in range(8):
        lib.ndsm_b200_profile_get(cls, ctypes.byref(cnt), ctypes.byref(tot))
        if cnt.value:
            avg_ms = tot.value / cnt.value
            kern[PROF_NAMES[cls]] = {"launches": cnt.value, "avg_ms": avg_ms, "total_ms": tot.value,
                                     "achieved_gbs": bytes_per_pt[cls] * npoints / (avg_ms * 1e-3) / 1e9}
    k0 = kern.get(PROF_NAMES[0], {"achieved_gbs": 0.0, "total_ms": 0.0})
    return {"bound": "hbm", "kernel": "k_relax3d<rhs=%s> colour pass" % ("0" if rhs_zero else "1"),
            "achieved": k0["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": k0["achieved_gbs"] / peak, "traffic": load_ncu_traffic().get(npoints), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": b0 * npoints, "share_of_step": k0["total_ms"] / (dev_ms if dev_ms else 1.0),
            "measured_in": "second pass of the same K steps with per-launch CUDA events (graphs off, one component at a time)",
            "kernels": kern}


def run_ours(args, rank, world):
    import torch
    from ndsm_b200 import load_library
    from ndsm_b200.ndsm import _options
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["NDSM_DEVICE"] = str(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load_library()
    if lib.ndsm_b200_device_count() <= 0:
        raise RuntimeError("bench.py: no CUDA device -- the product has no CPU fallback")
    cfg = Config(args, world)
    if cfg.name == "poisson_weak":
        return run_poisson(args, cfg, lib, rank, world, local, dist)

    nx, ny, nz = cfg.shape
    N = cfg.npoints
    x, y, z = cfg.mesh()
    b = cfg.field(x, y, z)
    nshape = np.array([nx, ny, nz, 3], dtype=np.intc)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ioptc, ropt = _options(lib, 10000, 1024, 1e-13, 1e-10, 5, cfg.mean, False)

    def barrier():
        if dist is not None:
            dist.barrier()

    def maxr(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    transport = "single GPU"
    if world == 1:
        # ---------------- device-resident arm: inputs already in HBM ------------------------------
        dB0 = torch.from_numpy(b).cuda()       # pristine boundary data
        dB = torch.empty_like(dB0)
        dA = torch.empty_like(dB0)
        npts_local = N

        def device_step():
            dA.zero_()                          # initial guess (reference passes zeros)
            dB.copy_(dB0)                       # B is overwritten by curl A every step
            torch.cuda.current_stream().synchronize()
            rc = lib.ndsm_b200_vector_solve_device(p(nshape), p(ioptc), p(ropt), p(x), p(y), p(z),
                                                   ctypes.c_void_p(dA.data_ptr()), ctypes.c_void_p(dB.data_ptr()))
            if rc != 0:
                raise RuntimeError("ndsm_b200_vector_solve_device returned %d" % rc)
        timed = "K x (zero A, restore B faces, ndsm_b200_vector_solve_device)"
    else:
        from ndsm_b200 import dist as ndist
        ndist.init_from_torch(local)
        transport = lib.ndsm_b200_dist_transport().decode()
        k0, k1 = ndist.slab_range(nz, world, rank)
        faces_h = ndist.extract_faces(b)
        faces_d = [torch.from_numpy(f).cuda() for f in faces_h]
        fptr = [f.data_ptr() for f in faces_d]
        dA = torch.empty((3, k1 - k0, ny, nx), dtype=torch.float64, device="cuda")
        dB = torch.empty_like(dA)
        npts_local = (k1 - k0) * ny * nx

        def device_step():
            rc, _, _, _ = ndist.vector_potential_rank(x, y, z, fptr, out=(dA.data_ptr(), dB.data_ptr()),
                                                      faces_on_device=True, mean=cfg.mean)
            if rc != 0:
                raise RuntimeError("ndsm_b200_vector_solve_rank returned %d" % rc)
        timed = "K x ndsm_b200_vector_solve_rank (six faces resident in HBM on every rank, z-slab outputs in HBM)"

    clocks = ClockSampler(local)
    clocks.start()   # started before the warm-up: spawning nvidia-smi stalls the driver for a moment
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize(); barrier()
    clocks.mark_begin()
    l0 = lib.ndsm_b200_launch_count()
    pb0, pm0 = lib.ndsm_b200_peer_bytes_sent(), lib.ndsm_b200_peer_messages_sent()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    upd_total, dev_ms = 0, 0.0
    tim = np.zeros(8)
    step_ms = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        device_step()
        step_ms.append((time.perf_counter() - ts) * 1e3)
        cyc = trace_cycles(lib, dist)
        upd_total += updates_from_cycles(cyc, N)
        lib.ndsm_b200_last_timing(p(tim))
        dev_ms += tim[6]
        stage = {"bc_ms": tim[2], "solve3d_ms": tim[3], "post_ms": tim[4]}
    ev1.record()
    torch.cuda.synchronize(); barrier()
    wall = maxr(time.perf_counter() - t0)
    clocks.mark_end()
    launches = lib.ndsm_b200_launch_count() - l0
    # NVLink traffic of the timed region: bytes / messages this rank stored into other GPUs' memory (peer transport)
    nvlink = None
    if world > 1:
        t = torch.tensor([lib.ndsm_b200_peer_bytes_sent() - pb0, lib.ndsm_b200_peer_messages_sent() - pm0],
                         dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        nvlink = {"bytes_per_step_all_ranks": float(t[0].item()) / args.steps,
                  "messages_per_step_all_ranks": float(t[1].item()) / args.steps,
                  "bytes_per_step_busiest_rank": float(tmax[0].item()) / args.steps,
                  "messages_per_step_busiest_rank": float(tmax[1].item()) / args.steps,
                  "counted": "remote stores of k_push (halo planes, replicated coarse levels, At faces, result pairs; "
                             "flag words excluded), counted on the host per launch / per graph replay"}
    clk = clocks.stop()
    value = upd_total / wall / 1e9
    nexact = [max([lib.ndsm_b200_trace_nexact(s, c) for c in range(lib.ndsm_b200_trace_ncycles(s))] or [0])
              for s in range(9)]

    # ---------------- multi-GPU parity, outside the timed region --------------------------------------
    # every rank repeats the solve on its own GPU alone (single-GPU path, no communication) and compares its
    # z-slab of the distributed result with it bit for bit
    check = {}
    if world > 1:
        Ad, Bd = dA.clone(), dB.clone()
        cyc_dist = [lib.ndsm_b200_trace_ncycles(s) for s in range(9)]
        fullB = torch.from_numpy(b).cuda()
        fullA = torch.zeros_like(fullB)
        rc = lib.ndsm_b200_vector_solve_device(p(nshape), p(ioptc), p(ropt), p(x), p(y), p(z),
                                               ctypes.c_void_p(fullA.data_ptr()), ctypes.c_void_p(fullB.data_ptr()))
        torch.cuda.synchronize()
        cyc_single = [lib.ndsm_b200_trace_ncycles(s) for s in range(9)]
        same = bool(rc == 0 and torch.equal(Ad, fullA[:, k0:k1]) and torch.equal(Bd, fullB[:, k0:k1]))
        dmax = float(max((Ad - fullA[:, k0:k1]).abs().max().item(), (Bd - fullB[:, k0:k1]).abs().max().item()))
        t = torch.tensor([1.0 if same else 0.0, -dmax], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        gathered = [None] * world
        dist.all_gather_object(gathered, {"rank": rank, "planes": [k0, k1], "v_cycles_AxAyAz": cyc_dist[6:],
                                          "single_gpu_v_cycles_AxAyAz": cyc_single[6:], "slab_equal": same})
        check = {"slab_bit_identical": bool(t[0].item() == 1.0), "max_abs_diff_vs_single_gpu": float(-t[1].item()),
                 "compared": "each rank's z-slab of A and B vs a single-GPU solve of the same problem on that rank's "
                             "own GPU (torch.equal)" + ("; mean metric: sums are order dependent, equality is not "
                                                        "expected, the difference must be rounding-level" if cfg.mean else ""),
                 "per_rank": gathered}
        del fullA, fullB, Ad, Bd
        torch.cuda.empty_cache()

    # Roofline pass: the same K steps again with per-launch CUDA events around every finest-level kernel
    # (event timing cannot be recorded inside the replayed CUDA graphs, so this pass launches directly).
    lib.ndsm_b200_profile_enable(1)
    prof_ms = 0.0
    for _ in range(args.steps):
        device_step()
        lib.ndsm_b200_last_timing(p(tim))
        prof_ms += tim[6]
    roofline = kernel_report(lib, int(lib.ndsm_b200_last_slab_points()) or npts_local, prof_ms, single_gpu=(world == 1))
    lib.ndsm_b200_profile_enable(0)

    # ---------------- end-to-end arm: host buffers, H2D and D2H inside the timed region -------------
    faces_bytes = 8 * 2 * (nx * ny + nx * nz + ny * nz) + 8 * (nx + ny + nz)
    e2e_upd, e2e_t = 0, 0.0
    if world == 1:   # the frozen reference-facing C ABI
        hA = torch.zeros(3 * N, dtype=torch.float64).pin_memory()
        hB = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hb0 = torch.from_numpy(b.reshape(-1))
        A_np, B_np = hA.numpy(), hB.numpy()
        api = "ndsm_vector_solve (frozen reference ABI), pinned host buffers"
    else:
        hA = torch.zeros(3 * npts_local, dtype=torch.float64).pin_memory()
        hB = torch.zeros(3 * npts_local, dtype=torch.float64).pin_memory()
        A_np, B_np = hA.numpy(), hB.numpy()
        fpin = [torch.from_numpy(f).pin_memory().numpy() for f in faces_h]
        api = "ndsm_b200_vector_solve_rank with host faces and host z-slab outputs (pinned)"
    d2h = 8 * 6 * N
    for i in range(args.warmup + args.steps):
        if world == 1:
            A_np[:] = 0.0
            hB.copy_(hb0)
        torch.cuda.synchronize(); barrier()
        t1 = time.perf_counter()
        if world == 1:
            rc = lib.ndsm_vector_solve(ctypes.c_size_t(3 * N), p(nshape), p(ioptc), p(ropt), p(x), p(y), p(z), p(A_np), p(B_np))
        else:
            fp = (ctypes.c_void_p * 6)(*[f.ctypes.data for f in fpin])
            rc = lib.ndsm_b200_vector_solve_rank(p(nshape), p(ioptc), p(ropt), p(x), p(y), p(z), fp, 0, p(A_np), p(B_np), 0)
        torch.cuda.synchronize(); barrier()
        dt = maxr(time.perf_counter() - t1)
        if rc != 0:
            raise RuntimeError("end-to-end solve returned %d" % rc)
        if i >= args.warmup:
            e2e_upd += updates_from_cycles(trace_cycles(lib, dist), N)
            e2e_t += dt
            lib.ndsm_b200_last_timing(p(tim))
            e2e_stage = {"in_ms": tim[1], "bc_ms": tim[2], "solve3d_ms": tim[3], "post_ms": tim[4], "d2h_ms": tim[5]}
    e2e = {"value": e2e_upd / e2e_t / 1e9, "unit": "Gpoint-updates/s", "h2d_bytes_per_step": faces_bytes * world,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_t / args.steps * 1e3, "stages_ms": e2e_stage,
           "host_memory": "pinned", "api": api}

    # analytic sanity of the last result (B against the exact field on the z = 0 face, rank 0's slab)
    if rank == 0:
        check["max_abs_B_error_on_z0_face"] = float(np.abs(B_np.reshape(3, -1, ny, nx)[:, 0] - b[:, 0]).max())
        check["max_abs_B_on_z0_face"] = float(np.abs(b[:, 0]).max())

    # ---------------- CPU baseline: bounded sample on the host cores (rank 0, N = 1 only) -----------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = physical_cores()
        v1, dt1, thr, desc = cpu_vcycle_sample(cfg, 1, threads)          # probe (also warms the page cache)
        ncyc = max(1, min(12, int(15.0 / max(dt1, 1e-3))))                # ~15 s of CPU work
        v, dt, thr, desc = cpu_vcycle_sample(cfg, ncyc, threads)
        cpu = {"value": v, "unit": "Gpoint-updates/s", "cores": thr, "kind": "port", "sample": desc, "seconds": dt}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Gpoint-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
            "scaling": cfg.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg.describe(),
            "details": {"v_cycles": {"chi": cyc[:6], "Ax": cyc[6], "Ay": cyc[7], "Az": cyc[8]},
                        "max_coarsest_iterations": nexact,
                        "decomposition": ("single GPU" if world == 1 else
                                          "%d ranks: every rank holds a z-slab of all three components (solved concurrently "
                                          "on three streams), chi faces distributed, transport: %s" % (world, transport)),
                        "steps_ms": step_ms, "timed_region": timed, "nvlink": nvlink},
            "time_to_vc_tol_ms": {"device_events": dev_ms / args.steps, "wall": wall / args.steps * 1e3, **stage,
                                  "torch_events_rank0": ev0.elapsed_time(ev1) / args.steps},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "check": check,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        lib.ndsm_b200_dist_finalize()
        dist.barrier()
        dist.destroy_process_group()


def run_poisson(args, cfg, lib, rank, world, local, dist):
    """BASELINE config 5: weak scaling of the scalar Poisson backend, 1025 x 1025 x (128 G + 1) on G GPUs."""
    import torch
    from ndsm_b200 import dist as ndist
    nx, ny, nz = cfg.shape
    N = cfg.npoints
    x, y, z = cfg.mesh()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    if world > 1:
        ndist.init_from_torch(local)
    transport = lib.ndsm_b200_dist_transport().decode() if world > 1 else "single GPU"
    k0, k1 = ndist.slab_range(nz, world, rank)
    uex, rhs = cfg.poisson_problem(k0, k1)
    d_rhs = torch.from_numpy(np.ascontiguousarray(rhs)).cuda()
    d_u = torch.zeros_like(d_rhs)
    h_rhs = torch.from_numpy(np.ascontiguousarray(rhs)).pin_memory()
    h_u = torch.zeros_like(h_rhs).pin_memory()
    npts_local = (k1 - k0) * ny * nx

    def barrier():
        if dist is not None:
            dist.barrier()

    def maxr(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {}

    def device_step():
        d_u.zero_()
        torch.cuda.current_stream().synchronize()
        ierr, du, nc = ndist.poisson_solve_rank(x, y, z, d_u.data_ptr(), d_rhs.data_ptr(), copt="NDDNDD")
        if ierr != 0:
            raise RuntimeError("ndsm_b200_poisson_solve_rank returned %d" % ierr)
        res["nc"], res["du"] = nc, du

    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize(); barrier()
    clocks.mark_begin()
    l0 = lib.ndsm_b200_launch_count()
    t0 = time.perf_counter()
    upd_total = 0
    step_ms = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        device_step()
        step_ms.append((time.perf_counter() - ts) * 1e3)
        upd_total += res["nc"] * 2 * 5 * N
    torch.cuda.synchronize(); barrier()
    wall = maxr(time.perf_counter() - t0)
    clocks.mark_end()
    launches = lib.ndsm_b200_launch_count() - l0
    clk = clocks.stop()
    value = upd_total / wall / 1e9
    err = float((d_u.cpu().numpy() - uex).__abs__().max())
    err = maxr(err)
    nexact = max([lib.ndsm_b200_trace_nexact(0, c) for c in range(lib.ndsm_b200_trace_ncycles(0))] or [0])

    lib.ndsm_b200_profile_enable(1)
    t1 = time.perf_counter()
    for _ in range(args.steps):
        device_step()
    torch.cuda.synchronize()
    prof_ms = (time.perf_counter() - t1) * 1e3
    roofline = kernel_report(lib, npts_local, prof_ms, rhs_zero=False, single_gpu=(world == 1))
    lib.ndsm_b200_profile_enable(0)

    # end to end: host slabs of u and rhs (pinned) -> device -> solve -> host
    e2e_t = 0.0
    for i in range(args.warmup + args.steps):
        h_u.zero_()
        torch.cuda.synchronize(); barrier()
        t1 = time.perf_counter()
        d_rhs.copy_(h_rhs, non_blocking=True)
        d_u.copy_(h_u, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        ierr, du, nc = ndist.poisson_solve_rank(x, y, z, d_u.data_ptr(), d_rhs.data_ptr(), copt="NDDNDD")
        h_u.copy_(d_u, non_blocking=True)
        torch.cuda.synchronize(); barrier()
        dt = maxr(time.perf_counter() - t1)
        if ierr != 0:
            raise RuntimeError("end-to-end poisson solve returned %d" % ierr)
        if i >= args.warmup:
            e2e_t += dt
    e2e = {"value": args.steps * res["nc"] * 10 * N / e2e_t / 1e9, "unit": "Gpoint-updates/s",
           "h2d_bytes_per_step": 16 * N, "d2h_bytes_per_step": 8 * N, "ms_per_step": e2e_t / args.steps * 1e3,
           "host_memory": "pinned", "api": "ndsm_b200_poisson_solve_rank around pinned host slabs of u and rhs"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = physical_cores()
        v1, dt1, thr, desc = cpu_vcycle_sample(cfg, 1, threads)
        ncyc = max(1, min(12, int(15.0 / max(dt1, 1e-3))))
        v, dt, thr, desc = cpu_vcycle_sample(cfg, ncyc, threads)
        cpu = {"value": v, "unit": "Gpoint-updates/s", "cores": thr, "kind": "port", "sample": desc, "seconds": dt}
    if rank == 0:
        h = x[1] - x[0]
        line = {
            "metric": METRIC, "value": value, "unit": "Gpoint-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
            "scaling": cfg.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg.describe(),
            "details": {"v_cycles": res["nc"], "du_last": res["du"], "max_coarsest_iterations": nexact,
                        "decomposition": "single GPU" if world == 1 else "%d z-slabs, transport: %s" % (world, transport),
                        "steps_ms": step_ms,
                        "timed_region": "K x (zero u, ndsm_b200_poisson_solve_rank on device slabs of u and rhs)"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "check": {"max_abs_error_vs_analytic": err, "h2": float(h * h), "error_over_h2": err / float(h * h)},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        lib.ndsm_b200_dist_finalize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="dipole513", choices=["dipole513", "charges1025x1025x257", "poisson_weak"])
    ap.add_argument("--size", dest="n", type=int, default=513, help="mesh points per dimension (dipole config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
