"""How fast is the finest-level colour pass when the level fits in L2?  (513 x 513 x nz slabs of growing nz)
Decides whether temporal blocking through the 126 MB L2 can pay: prints ns per plane per colour pass."""
import ctypes, os, sys, time
os.environ.setdefault('NDSM_B200_HANDLE_RHS0', '0')
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndsm_b200 import load_library
from ndsm_b200.mg import MGHandle
lib = load_library()
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 513
rng = np.random.default_rng(0)
NZ = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else (513, 257, 129, 65, 49, 33, 25, 17, 9)
for nz in NZ:
    x = np.linspace(0, 1, nx); dx = x[1] - x[0]
    mesh = [x, np.arange(nx) * dx, np.arange(nz) * dx]
    h = MGHandle(mesh, "NDDNDD")
    h.put(0, 0, rng.standard_normal((nz, nx, nx)))
    h.relax(0, max(20, 40000 // nz))
    # (a) back-to-back launches, no events in between: wall time of a long queue / number of passes
    n = max(50, 40000 // nz)
    t0 = time.perf_counter(); h.relax(0, n); wall = (time.perf_counter() - t0) / (2 * n) * 1e3
    best = 1e9
    for rep in range(3):
        lib.ndsm_b200_profile_enable(1)
        h.relax(0, max(20, 10000 // nz))
        cnt = ctypes.c_ulonglong(); ms = ctypes.c_double()
        lib.ndsm_b200_profile_get(0, ctypes.byref(cnt), ctypes.byref(ms))
        lib.ndsm_b200_profile_enable(0)
        best = min(best, ms.value / max(cnt.value, 1))
    mb = nx * nx * nz * 8 / 1e6
    print("nz %4d  u %.0f MB  pass %.2f us (events) %.2f us (queue) -> %.1f ns/plane  (%.0f GB/s algorithmic u-only)" % (nz, mb, best * 1e3, wall * 1e3, min(best, wall) * 1e6 / nz, mb / 1e3 / (min(best, wall) / 1e3)), flush=True)
    h.close()
