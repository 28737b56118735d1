"""ndsm_b200 -- B200 (sm_100a) drop-in for the multigrid vector-potential solve of sag2021/ndsm.

The product is the C-ABI shared library ``ndsm_b200/lib/ndsmf.so`` (CUDA kernels + host driver).
This package only holds the host-side mirror of the reference's Python interface:

* :func:`ndsm_b200.ndsm.vector_potential` -- same signature/semantics as the reference's
  ``ndsm.vector_potential`` (ndsm.py:66-210).
* :class:`ndsm_b200.mg.MGHandle` -- ctypes view of the MG_HANDLE operator seam
  (ndsm_multigrid_core.f90:86-136) used by the parity tests.
"""
from .lib_loader import LIB_PATH, load_library  # noqa: F401
from .ndsm import vector_potential  # noqa: F401

__all__ = ["vector_potential", "load_library", "LIB_PATH"]
