"""Micro-benchmark of the finest-level transfer kernels (restriction 513^3 -> 256^3, prolongation back) through the
MG_HANDLE seam, for the tuning variants selected by environment variables read per launch."""
import ctypes
import os
import sys
import time

import numpy as np

os.environ.setdefault("NDSM_B200_HANDLE_RHS0", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndsm_b200.mg import MGHandle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
x = np.linspace(0, 1, n)
h = MGHandle([x, x.copy(), x.copy()], "NDDNDD")
rng = np.random.default_rng(0)
r = rng.standard_normal((n, n, n))
h.put(h.R, 0, r)
h.put(h.U, 0, r)
h.put(h.U, 1, rng.standard_normal(h.shape(1)[::-1]))
lib = h.lib


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


N = n ** 3
for mb in ("2", "3", "4"):
    os.environ["NDSM_B200_RD_MINB"] = mb
    ms = timed(lambda: lib.ndsm_b200_mg_restrict(h.h, 0))
    print("restrict_direct MINB=%s  %.4f ms  (%.0f GB/s of 9 B/pt)" % (mb, ms, 9.0 * N / ms / 1e6), flush=True)
os.environ.pop("NDSM_B200_RD_MINB", None)
for npv in ("2", "4", "6", "8"):
    os.environ["NDSM_B200_IZ_NP"] = npv
    ms = timed(lambda: lib.ndsm_b200_mg_interp_add(h.h, 1))
    print("interp_add_zt NP=%s  %.4f ms  (%.0f GB/s of 17 B/pt)" % (npv, ms, 17.0 * N / ms / 1e6), flush=True)
os.environ.pop("NDSM_B200_IZ_NP", None)
ms = timed(lambda: lib.ndsm_b200_mg_residual(h.h, 0))
print("residual3d  %.4f ms  (%.0f GB/s of 16 B/pt)" % (ms, 16.0 * N / ms / 1e6))
ms = timed(lambda: lib.ndsm_b200_mg_relax(h.h, 0, 1))
print("relax sweep (2 passes) %.4f ms  (%.0f GB/s of 16 B/pt)" % (ms, 16.0 * N / ms / 1e6))
