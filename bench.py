#!/usr/bin/env python
"""bench.py -- headline benchmark of the NDSM vector-potential hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 513]

A "step" is one complete vector-potential solve (BC setup: 6 chi solves -> 3 multigrid solves to
vc_tol -> flux-balance fields -> curl) of the synthetic sub-surface dipole on an n^3 mesh (default
513^3, the configuration BASELINE.json's metric is quoted on), FP64, default tolerances.

value   : fine-grid Gpoint-updates/s of the device-resident solve (inputs already in HBM), i.e.
          sum_c(V-cycles_c * 2*ms * nx*ny*nz) / time  (SURVEY.md 8d), whole job over all N GPUs.
e2e     : the same metric through the frozen reference-facing C ABI `ndsm_vector_solve` with HOST
          buffers (pinned), H2D of the faces and D2H of A and B inside the timed region.
roofline: the dominant kernel (k_relax3d colour pass on the finest level, rhs == 0) timed live with
          CUDA events on the library's stream; algorithmic bytes = 8 B per fine-grid point per launch.
cpu_baseline / --impl reference: the CPU oracle (restatement of the reference's OpenMP path; the
          Fortran reference cannot be compiled in this image) on a bounded sample of the same workload.
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


def workload_name(n):
    return "dipole-%d^3 vector_potential, default options (max metric, vc_tol=1e-10, ex_tol=1e-13, ms=5)" % n


def workload(n, faces_only=True):
    from ndsm_b200 import synthetic
    x, y, z = synthetic.mesh(n)
    b = synthetic.dipole(x, y, z, faces_only=faces_only)
    return x, y, z, b


def updates_from_trace(lib, n_points, ms=5, dist=None):
    """sum over Ax,Ay,Az of V-cycles * 2*ms * N.  Az always uses ms=5 (reference quirk).
    With >= 3 ranks every rank only traces the component its group solved: take the max over ranks."""
    cyc = [lib.ndsm_b200_trace_ncycles(s) for s in range(9)]
    if dist is not None:
        import torch
        t = torch.tensor(cyc, dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cyc = [int(v) for v in t.tolist()]
    upd = sum(cyc[6 + c] * 2 * (5 if c == 2 else ms) * n_points for c in range(3))
    return upd, cyc


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle): bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_sample(n, ncycles=1):
    """`ncycles` V-cycles of the Ax-type solve (copt NDDNDD, rhs = 0, Dirichlet data on the y/z faces) on
    the n^3 mesh with the oracle.  Returns (Gpoint-updates/s, seconds, threads, description)."""
    from oracle import pyoracle as O
    from ndsm_b200 import synthetic
    x, y, z = synthetic.mesh(n)
    u0 = np.zeros((n, n, n))
    # smooth Dirichlet data on the four D faces (values do not affect the sweep cost)
    yy, xx = np.meshgrid(y, x, indexing="ij")
    u0[0, :, :] = np.sin(np.pi * xx) * np.cos(np.pi * yy)
    u0[-1, :, :] = 0.1 * u0[0, :, :]
    u0[:, 0, :] = u0[0, 0, :][None, :]
    u0[:, -1, :] = u0[0, -1, :][None, :]
    mg = O.OracleMG([x, y, z], "NDDNDD", ms=5)
    rhs = np.zeros_like(u0)
    mg.load(u0, rhs)
    t0 = time.perf_counter()
    nc = 0
    for _ in range(ncycles):       # solve_poisson_bvp's loop body: V-cycle + update_u
        mg.v_cycle()
        O.update_u(mg.u(0), u0)
        nc += 1
    dt = time.perf_counter() - t0
    mg.close()
    upd = nc * 2 * 5 * n ** 3
    return upd / dt / 1e9, dt, O.num_threads(), "%d V-cycle(s) of the Ax solve (copt NDDNDD, ms=5) on %d^3" % (nc, n)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The Fortran reference cannot be
    built in this image (no gfortran), so this times the oracle port with all host threads."""
    if rank != 0:
        return
    n = args.n
    # keep the whole run within a few minutes: one V-cycle per step, mesh reduced if a step is too slow
    val, dt, thr, desc = cpu_sample(min(n, 257), 1)
    if dt * 8 * (n / min(n, 257)) ** 3 < 25.0:
        sample_n = n
    else:
        sample_n = min(n, 257)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, dt, thr, desc = cpu_sample(sample_n, 1)
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = float(np.sum([10 * sample_n ** 3 for _ in vals]) / np.sum(times) / 1e9)
    line = {
        "impl": "reference", "metric": "fine-grid Gpoint-updates/s (time-to-vc_tol in ms_per_step)", "value": value,
        "unit": "Gpoint-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "sample": desc},
        "cpu_baseline": {"value": value, "unit": "Gpoint-updates/s", "cores": thr, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "Gpoint-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def kernel_report(lib, npoints, dev_ms):
    """Per-kernel-class CUDA-event timings gathered by the library (finest level of this rank's slab)."""
    cnt, tot = ctypes.c_ulonglong(0), ctypes.c_double(0)
    peak, peak_src = measured_peak()
    kern = {}
    names = ["k_relax3d<rhs=0> colour pass", "k_residual3d", "k_restrict_direct", "k_interp_add_zt",
             "update_u (k_diff_partial+final)", "halo exchange (one NCCL group)", "levels >= 2 of one V-cycle",
             "level 1 of one V-cycle (2 brackets per cycle)"]
    bytes_per_pt = [8.0, 16.0, 9.0, 17.0, 24.0, 0.0, 0.0, 0.0]  # SURVEY 8d / DESIGN.md (level 0, rhs == 0)
    for cls in range(8):
        lib.ndsm_b200_profile_get(cls, ctypes.byref(cnt), ctypes.byref(tot))
        if cnt.value:
            avg_ms = tot.value / cnt.value
            kern[names[cls]] = {"launches": cnt.value, "avg_ms": avg_ms, "total_ms": tot.value,
                                "achieved_gbs": bytes_per_pt[cls] * npoints / (avg_ms * 1e-3) / 1e9}
    k0 = kern.get(names[0], {"achieved_gbs": 0.0, "total_ms": 0.0})
    return {"bound": "hbm", "kernel": names[0], "achieved": k0["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": k0["achieved_gbs"] / peak, "traffic": NCU_TRAFFIC_BYTES.get(npoints), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": 8.0 * npoints, "share_of_step": k0["total_ms"] / (dev_ms if dev_ms else 1.0),
            "measured_in": "second pass of the same K steps with per-launch CUDA events (graphs off)",
            "kernels": kern}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the finest-level colour pass from the committed
# `ncu --set full` capture (profiles/), keyed by points per launch
NCU_TRAFFIC_BYTES = {513 ** 3: 562.7e6 + 497.0e6}  # profiles/r01_ncu_full_finest_level_kernels_513.json


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

    n = args.n
    N = n ** 3
    x, y, z, b = workload(n)
    nshape = np.array([n, n, n, 3], dtype=np.intc)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    ioptc, ropt = _options(lib, 10000, 1024, 1e-13, 1e-10, 5, False, False)

    def barrier():
        if dist is not None:
            dist.barrier()

    def maxr(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
        k0, k1 = ndist.slab_range(n, world, rank)
        faces_h = ndist.extract_faces(b)
        faces_d = [torch.from_numpy(f).cuda() for f in faces_h]
        fptr = [f.data_ptr() for f in faces_d]
        dA = torch.empty((3, k1 - k0, n, n), dtype=torch.float64, device="cuda")
        dB = torch.empty_like(dA)
        npts_local = (k1 - k0) * n * n

        def device_step():
            rc, _, _, _ = ndist.vector_potential_rank(x, y, z, fptr, out=(dA.data_ptr(), dB.data_ptr()),
                                                      faces_on_device=True)
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
        upd, cyc = updates_from_trace(lib, N, dist=dist)
        upd_total += upd
        lib.ndsm_b200_last_timing(p(tim))
        dev_ms += tim[6]
        stage = {"bc_ms": tim[2], "solve3d_ms": tim[3], "post_ms": tim[4]}
    ev1.record()
    torch.cuda.synchronize(); barrier()
    wall = maxr(time.perf_counter() - t0)
    clocks.mark_end()
    launches = lib.ndsm_b200_launch_count() - l0
    clk = clocks.stop()
    value = upd_total / wall / 1e9

    # Roofline pass: the same K steps again with per-launch CUDA events around every finest-level kernel
    # (event timing cannot be recorded inside the replayed CUDA graphs, so this pass launches directly).
    lib.ndsm_b200_profile_enable(1)
    prof_ms = 0.0
    for _ in range(args.steps):
        device_step()
        lib.ndsm_b200_last_timing(p(tim))
        prof_ms += tim[6]
    roofline = kernel_report(lib, int(lib.ndsm_b200_last_slab_points()) or npts_local, prof_ms)
    lib.ndsm_b200_profile_enable(0)

    # ---------------- end-to-end arm: host buffers, H2D and D2H inside the timed region -------------
    faces_bytes = 8 * 6 * n * n + 8 * 3 * n
    e2e_upd, e2e_t = 0, 0.0
    if world == 1:   # the frozen reference-facing C ABI
        hA = torch.zeros(3 * N, dtype=torch.float64).pin_memory()
        hB = torch.empty(3 * N, dtype=torch.float64).pin_memory()
        hb0 = torch.from_numpy(b.reshape(-1))
        A_np, B_np = hA.numpy(), hB.numpy()
        d2h = 8 * 6 * N
        api = "ndsm_vector_solve (frozen reference ABI), pinned host buffers"
    else:
        hA = torch.zeros(3 * npts_local, dtype=torch.float64).pin_memory()
        hB = torch.zeros(3 * npts_local, dtype=torch.float64).pin_memory()
        A_np, B_np = hA.numpy(), hB.numpy()
        fpin = [torch.from_numpy(f).pin_memory().numpy() for f in faces_h]
        d2h = 8 * 6 * N
        api = "ndsm_b200_vector_solve_rank with host faces and host z-slab outputs (pinned)"
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
            e2e_upd += updates_from_trace(lib, N, dist=dist)[0]
            e2e_t += dt
            lib.ndsm_b200_last_timing(p(tim))
            e2e_stage = {"in_ms": tim[1], "bc_ms": tim[2], "solve3d_ms": tim[3], "post_ms": tim[4], "d2h_ms": tim[5]}
    e2e = {"value": e2e_upd / e2e_t / 1e9, "unit": "Gpoint-updates/s", "h2d_bytes_per_step": faces_bytes * world,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_t / args.steps * 1e3, "stages_ms": e2e_stage,
           "host_memory": "pinned", "api": api}

    # analytic sanity of the last result (B against the exact dipole field on the z = 0 face, rank 0's slab)
    err = float(np.abs(B_np.reshape(3, -1, n, n)[:, 0] - b[:, 0]).max()) if rank == 0 else None

    # ---------------- CPU baseline: bounded sample on the host cores (rank 0, N = 1 only) -----------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sample_n = n if n <= 257 else 257
        v, dt, thr, desc = cpu_sample(sample_n, 1)
        if n > sample_n and dt * (n / sample_n) ** 3 < 30.0:
            v, dt, thr, desc = cpu_sample(n, 1)
        cpu = {"value": v, "unit": "Gpoint-updates/s", "cores": thr, "kind": "port", "sample": desc, "seconds": dt}

    if rank == 0:
        line = {
            "metric": "fine-grid Gpoint-updates/s (time-to-vc_tol in ms_per_step)", "value": value,
            "unit": "Gpoint-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "l2": "inputs larger than L2 (each %d^3 fp64 array = %.2f GB)" % (n, 8 * N / 1e9),
                       "v_cycles": {"chi": cyc[:6], "Ax": cyc[6], "Ay": cyc[7], "Az": cyc[8]},
                       "decomposition": ("single GPU" if world == 1 else
                                         "2 ranks: z-slabs, communication-avoiding halo exchange over NCCL" if world == 2 else
                                         "%d ranks: three component groups (Ax|Ay|Az) x z-slabs inside each group, "
                                         "NCCL halo exchange + all-to-all of A before the curl" % world),
                       "steps_ms": step_ms,
                       "timed_region": timed},
            "time_to_vc_tol_ms": {"device_events": dev_ms / args.steps, "wall": wall / args.steps * 1e3, **stage,
                                  "torch_events_rank0": ev0.elapsed_time(ev1) / args.steps},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "check": {"max_abs_B_error_on_z0_face": err},
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
    ap.add_argument("--size", dest="n", type=int, default=513, help="mesh points per dimension")
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
