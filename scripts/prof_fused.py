"""Profiling target: the fused residual+restriction of the finest level pair (and the unfused pair) at n^3."""
import os
os.environ.setdefault("NDSM_B200_HANDLE_RHS0", "0")
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndsm_b200.mg import MGHandle  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
x = np.linspace(0, 1, n)
h = MGHandle([x, x.copy(), x.copy()], "NDDNDD")
r = np.random.default_rng(0).standard_normal((n, n, n))
h.put(h.U, 0, r)
import ctypes
f = ctypes.c_int(0)
for _ in range(3):
    h.lib.ndsm_b200_mg_residual_restrict(h.h, 0, ctypes.byref(f))
print("fused", f.value)
