"""Capacity / correctness check of the remaining BASELINE configs on one GPU:
  config 3: non-cubic active-region-like box 1025x1025x257 (two sub-surface charges), device-resident entry
  config 5: scalar Poisson backend, 1025x1025x129 (the G = 1 point of the weak-scaling series), NDDNDD, analytic rhs
Prints one JSON line per config (times, V-cycles, errors against the analytic field)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ndsm_b200 import load_library, synthetic  # noqa: E402
from ndsm_b200.ndsm import _options, read_timing  # noqa: E402

lib = load_library()
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def config3(nx=1025, ny=1025, nz=257):
    x, y, z = synthetic.mesh(nx, ny, nz)
    b = synthetic.charges(x, y, z, faces_only=True)
    dB = torch.from_numpy(b).cuda()
    exact_z0 = b[:, 0].copy()
    del b
    dA = torch.zeros_like(dB)
    nshape = np.array([nx, ny, nz, 3], dtype=np.intc)
    ioptc, ropt = _options(lib, 10000, 1024, 1e-13, 1e-10, 5, True, False)  # mean metric
    out = []
    for rep in range(2):
        dA.zero_()
        if rep:
            dB.copy_(dB0)
        else:
            dB0 = dB.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.ndsm_b200_vector_solve_device(p(nshape), p(ioptc), p(ropt), p(x), p(y), p(z), ctypes.c_void_p(dA.data_ptr()),
                                               ctypes.c_void_p(dB.data_ptr()))
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
    cyc = [lib.ndsm_b200_trace_ncycles(s) for s in range(9)]
    nex = [max(lib.ndsm_b200_trace_nexact(s, c) for c in range(cyc[s])) for s in range(9)]
    Bz0 = dB[:, 0].cpu().numpy()
    err = np.abs(Bz0 - exact_z0)
    print(json.dumps({"config": "3: charges magnetogram %dx%dx%d, mean metric, 1 GPU" % (nx, ny, nz), "rc": rc, "ms": out,
                      "stages": read_timing(lib), "v_cycles": cyc, "max_coarsest_iterations": nex,
                      "B_err_z0_face_max": float(err.max()), "B_err_z0_face_mean": float(err.mean()),
                      "B_max": float(np.abs(exact_z0).max())}), flush=True)


def config5(nx=1025, ny=1025, nz=129):
    x = np.linspace(0, 1, nx)
    h = x[1] - x[0]
    y = np.arange(ny) * h
    z = np.arange(nz) * h
    Ly, Lz = y[-1], z[-1]
    # u = cos(pi x) sin(pi y / Ly) sin(pi z / Lz): Neumann in x, Dirichlet 0 in y and z
    cx, sy, sz = np.cos(np.pi * x), np.sin(np.pi * y / Ly), np.sin(np.pi * z / Lz)
    uex = sz[:, None, None] * sy[None, :, None] * cx[None, None, :]
    lam = -(np.pi ** 2) * (1.0 + 1.0 / Ly ** 2 + 1.0 / Lz ** 2)
    rhs = lam * uex
    u = np.zeros_like(uex)
    nshape = np.array([nx, ny, nz], dtype=np.intc)
    du = ctypes.c_double(0)
    nc = ctypes.c_int(0)
    t0 = time.perf_counter()
    rc = lib.ndsm_b200_poisson_solve(3, p(nshape), b"NDDNDD", 5, 1024, 10000, 1, 1e-10, 1e-13, p(x), p(y), p(z), p(u), p(rhs),
                                     ctypes.byref(du), ctypes.byref(nc))
    dt = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"config": "5: scalar Poisson %dx%dx%d NDDNDD (G=1 of the weak-scaling series), host arrays" % (nx, ny, nz),
                      "rc": rc, "ms_including_h2d_d2h": dt, "v_cycles": nc.value, "du_last": du.value,
                      "max_abs_error_vs_analytic": float(np.abs(u - uex).max()), "h2": float(h * h)}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    if which in ("3", "both"):
        config3()
    if which in ("5", "both"):
        config5()
