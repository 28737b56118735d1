"""Profiling target: restriction and prolongation of the finest level pair of an n^3 hierarchy (MG_HANDLE seam)."""
import os
os.environ.setdefault("NDSM_B200_HANDLE_RHS0", "0")  # rhs == 0 specialisations, as on the finest level of the solves
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndsm_b200.mg import MGHandle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
x = np.linspace(0, 1, n)
h = MGHandle([x, x.copy(), x.copy()], "NDDNDD")
rng = np.random.default_rng(0)
r = rng.standard_normal((n, n, n))
h.put(h.R, 0, r)
h.put(h.U, 0, r)
nc = h.shape(1)
h.put(h.U, 1, rng.standard_normal(nc[::-1]))
for _ in range(3):
    h.lib.ndsm_b200_mg_restrict(h.h, 0)
    h.lib.ndsm_b200_mg_interp_add(h.h, 1)
    h.lib.ndsm_b200_mg_residual(h.h, 0)
    h.lib.ndsm_b200_mg_relax(h.h, 0, 1)
    h.update_u(r, r)
print("done")
