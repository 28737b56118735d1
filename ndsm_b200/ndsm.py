"""Host-side mirror of the reference's Python interface (ndsm.py:66-210).

``vector_potential`` keeps the reference's name, arguments, defaults and return value; it fills the
two 16-slot option vectors through the library's index getters exactly like the reference does
(ndsm.py:155-199) and calls the frozen C entry ``ndsm_vector_solve`` (ndsm_python_wrapper.f90:56).
The reference's own unmodified ndsm.py works against the same library; this
mirror exists so that tests and benchmarks do not need the reference tree at run time.

Differences from the reference wrapper: ``libpath=None`` resolves to the in-tree
``ndsm_b200/lib/ndsmf.so`` instead of walking ``sys.path``; ``trace=True`` additionally returns the
per-solve V-cycle history recorded by the library.
"""
import ctypes

import numpy as np

from .lib_loader import _declare, load_library

SOLVE_NAMES = ("chi1", "chi2", "chi3", "chi4", "chi5", "chi6", "Ax", "Ay", "Az")


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _options(lib, niterex_max, ncycles_max, ex_tol, vc_tol, ms, mean, debug, flxcrl=0):
    n = lib.get_iopt_len()
    ioptc = np.zeros(n, dtype=np.intc)
    ropt = np.zeros(n, dtype=np.float64)
    ioptc[lib.get_iopt_ms()] = ms
    ioptc[lib.get_iopt_ncycles()] = ncycles_max
    ioptc[lib.get_iopt_iopt_nmaxex()] = niterex_max
    ropt[lib.get_ropt_vtol()] = vc_tol
    ropt[lib.get_ropt_ctol()] = ex_tol
    ioptc[lib.get_iopt_debug()] = lib.get_iopt_true() if debug else lib.get_iopt_false()
    ioptc[lib.get_iopt_dumax()] = lib.get_iopt_false() if mean else lib.get_iopt_true()
    ioptc[4] = flxcrl  # IOPT_FLXCRL (ndsm_vector_potential.f90:45); no getter exists in the reference
    return ioptc, ropt


def read_trace(lib):
    """V-cycle history of the last solve: {name: {"du": [...], "nexact": [...]}}."""
    out = {}
    for s, name in enumerate(SOLVE_NAMES):
        nc = lib.ndsm_b200_trace_ncycles(s)
        out[name] = {"du": [lib.ndsm_b200_trace_du(s, c) for c in range(nc)],
                     "nexact": [lib.ndsm_b200_trace_nexact(s, c) for c in range(nc)]}
    return out


def read_timing(lib):
    t = np.zeros(8)
    lib.ndsm_b200_last_timing(_ptr(t))
    keys = ("ms_total", "ms_in", "ms_bc", "ms_solve3d", "ms_post", "ms_out", "ms_device", "launches")
    return dict(zip(keys, t.tolist()))


def vector_potential(x, y, z, b, niterex_max=10000, ncycles_max=1024, ex_tol=1e-13, vc_tol=1e-10, ms=5,
                     mean=False, libname="ndsmf.so", libpath=None, debug=False, flxcrl=0, A0=None, trace=False):
    """Vector potential A and B = curl A of the potential field matching the normal component of ``b``
    on the six faces of the box (reference: ndsm.py:66).

    x, y, z : (nx,), (ny,), (nz,) uniform mesh vectors.  b : (3, nz, ny, nx); only boundary normal
    components are read.  Returns ``(ierr, A[3,nz,ny,nx], B[3,nz,ny,nx])`` (plus the trace dict when
    ``trace=True``).  ``A0`` optionally supplies the initial guess (the reference always passes zeros).
    """
    if libpath is None:
        lib = load_library()
    else:
        lib = ctypes.CDLL(libpath)
        _declare(lib)
    b = np.asarray(b, dtype=np.float64)
    if b.ndim != 4 or b.shape[0] != 3:
        raise ValueError("b must have shape (3, nz, ny, nx)")
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    nshape = np.array(b.shape[::-1], dtype=np.intc)  # Fortran order [nx,ny,nz,3] (ndsm.py:162)
    if (x.size, y.size, z.size) != (nshape[0], nshape[1], nshape[2]):
        raise ValueError("mesh vectors do not match b.shape")
    ioptc, ropt = _options(lib, niterex_max, ncycles_max, ex_tol, vc_tol, ms, mean, debug, flxcrl)
    Apot = np.zeros(b.size, dtype=np.float64) if A0 is None else np.array(A0, dtype=np.float64).ravel().copy()
    bflat = b.flatten()  # copy, like the reference (ndsm.py:204)
    ierr = lib.ndsm_vector_solve(ctypes.c_size_t(b.size), _ptr(nshape), _ptr(ioptc), _ptr(ropt), _ptr(x), _ptr(y),
                                 _ptr(z), _ptr(Apot), _ptr(bflat))
    shp = tuple(int(v) for v in nshape[::-1])
    res = (ierr, Apot.reshape(shp), bflat.reshape(shp))
    if trace:
        info = read_trace(lib)
        info["timing"] = read_timing(lib)
        info["seconds"] = float(ropt[lib.get_ropt_tim()])
        info["ioptc"] = ioptc.copy()
        return res + (info,)
    return res
