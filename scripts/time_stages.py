"""Stage timings of the device-resident solve: usage time_stages.py n [reps]  (env NDSM_B200_GRAPH=0/1)"""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ndsm_b200 import load_library, synthetic  # noqa: E402
from ndsm_b200.ndsm import _options, read_timing  # noqa: E402

n = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
lib = load_library()
x, y, z = synthetic.mesh(n)
b = synthetic.dipole(x, y, z, faces_only=True)
dB0 = torch.from_numpy(b).cuda()
dB = torch.empty_like(dB0)
dA = torch.zeros_like(dB0)
nshape = np.array([n, n, n, 3], dtype=np.intc)
ioptc, ropt = _options(lib, 10000, 1024, 1e-13, 1e-10, 5, False, False)
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
for r in range(reps):
    dA.zero_(); dB.copy_(dB0); torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = lib.ndsm_b200_vector_solve_device(p(nshape), p(ioptc), p(ropt), p(x), p(y), p(z), ctypes.c_void_p(dA.data_ptr()),
                                           ctypes.c_void_p(dB.data_ptr()))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) * 1e3
    t = read_timing(lib)
    cyc = [lib.ndsm_b200_trace_ncycles(s) for s in range(9)]
    print("n=%d graph=%s rep=%d rc=%d wall %.1f ms | bc %.1f solve3d %.1f post %.1f dev %.1f | launches %d | cycles %s" % (
        n, os.environ.get("NDSM_B200_GRAPH", "1"), r, rc, dt, t["ms_bc"], t["ms_solve3d"], t["ms_post"], t["ms_device"],
        t["launches"], cyc), flush=True)
