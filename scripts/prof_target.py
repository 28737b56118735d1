"""Short profiling target: one device-resident vector-potential solve on an n^3 dipole.
usage: prof_target.py n [max_vcycles]"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ndsm_b200 import load_library, synthetic  # noqa: E402
from ndsm_b200.ndsm import _options  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 257
ncyc = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
lib = load_library()
x, y, z = synthetic.mesh(n)
b = synthetic.dipole(x, y, z, faces_only=True)
dB = torch.from_numpy(b).cuda()
dA = torch.zeros_like(dB)
nshape = np.array([n, n, n, 3], dtype=np.intc)
ioptc, ropt = _options(lib, 10000, ncyc, 1e-13, 1e-10, 5, False, False)
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
torch.cuda.synchronize()
rc = lib.ndsm_b200_vector_solve_device(p(nshape), p(ioptc), p(ropt), p(x), p(y), p(z), ctypes.c_void_p(dA.data_ptr()),
                                       ctypes.c_void_p(dB.data_ptr()))
torch.cuda.synchronize()
print("rc", rc, "seconds", ropt[2], "launches", lib.ndsm_b200_launch_count())
sys.exit(0 if rc in (0, 1) else 1)
