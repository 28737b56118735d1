"""Tuning sweep of the smoother's z-chunk length on slab-shaped levels (what one rank of a z-slab decomposition
smooths).  One process; NDSM_B200_ZCHUNK_FORCE is read by the library on every launch."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndsm_b200.mg import MGHandle  # noqa: E402


def case(nx, ny, nz, rhs):
    os.environ["NDSM_B200_HANDLE_RHS0"] = "1" if rhs else "0"
    x = np.linspace(0, 1, nx)
    h = x[1] - x[0]
    m = MGHandle([x, np.arange(ny) * h, np.arange(nz) * h], "NDDNDD", ngrids=1)
    lib = m.lib
    cnt, tot = ctypes.c_ulonglong(0), ctypes.c_double(0)
    out = []
    for zc in ["auto", 4, 6, 8, 10, 12, 13, 16, 19, 22, 26, 32]:
        if zc == "auto":
            os.environ.pop("NDSM_B200_ZCHUNK_FORCE", None)
        elif zc > nz:
            continue
        else:
            os.environ["NDSM_B200_ZCHUNK_FORCE"] = str(zc)
        lib.ndsm_b200_profile_enable(0)
        m.relax(0, 3)
        lib.ndsm_b200_profile_enable(1)
        m.relax(0, 20)
        lib.ndsm_b200_profile_get(0, ctypes.byref(cnt), ctypes.byref(tot))
        us = tot.value / max(cnt.value, 1) * 1e3
        out.append((zc, us))
    lib.ndsm_b200_profile_enable(0)
    os.environ.pop("NDSM_B200_ZCHUNK_FORCE", None)
    m.close()
    npts = nx * ny * nz
    bpp = 12.0 if rhs else 8.0
    best = min(out, key=lambda t: t[1])
    print("%4dx%4dx%3d rhs=%d  " % (nx, ny, nz, rhs) + "  ".join("%s:%.1f" % (z, u) for z, u in out) +
          "   best zc=%s %.1f us = %.0f GB/s" % (best[0], best[1], bpp * npts / (best[1] * 1e-6) / 1e9), flush=True)


if __name__ == "__main__":
    for c in [(513, 513, 64, 0), (513, 513, 70, 0), (513, 513, 76, 0), (513, 513, 128, 0), (513, 513, 136, 0),
              (513, 513, 256, 0), (513, 513, 513, 0), (256, 256, 32, 1), (256, 256, 40, 1), (256, 256, 64, 1),
              (256, 256, 128, 1), (256, 256, 256, 1), (1025, 1025, 32, 0), (1025, 1025, 44, 0), (1025, 1025, 129, 1)]:
        case(*c)
