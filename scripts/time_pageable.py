"""End-to-end time of ndsm_vector_solve with ordinary (pageable) numpy buffers vs pinned ones."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ndsm_b200 import load_library, synthetic
from ndsm_b200.ndsm import _options, read_timing
n = int(sys.argv[1]) if len(sys.argv) > 1 else 513
lib = load_library()
x, y, z = synthetic.mesh(n)
b = synthetic.dipole(x, y, z, faces_only=True)
N = n ** 3
nshape = np.array([n, n, n, 3], dtype=np.intc)
ioptc, ropt = _options(lib, 10000, 1024, 1e-13, 1e-10, 5, False, False)
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
for kind in ("pageable", "pinned"):
    if kind == "pinned":
        A = torch.zeros(3 * N, dtype=torch.float64).pin_memory().numpy()
        B = torch.zeros(3 * N, dtype=torch.float64).pin_memory().numpy()
    else:
        A = np.zeros(3 * N); B = np.zeros(3 * N)
    for rep in range(3):
        A[:] = 0.0; B[:] = b.reshape(-1)
        t0 = time.perf_counter()
        rc = lib.ndsm_vector_solve(ctypes.c_size_t(3 * N), p(nshape), p(ioptc), p(ropt), p(x), p(y), p(z), p(A), p(B))
        dt = (time.perf_counter() - t0) * 1e3
        t = read_timing(lib)
        print("%s rep %d rc %d total %.1f ms | in %.1f bc %.1f solve %.1f post %.1f d2h %.1f" % (kind, rep, rc, dt, t["ms_in"], t["ms_bc"], t["ms_solve3d"], t["ms_post"], t["ms_out"]), flush=True)
