"""ctypes wrapper of the CPU oracle (oracle/_build/libndsm_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under ndsm_b200/ may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libndsm_oracle.so")
_LIB = None
c = ctypes
vp = c.c_void_p
i64 = c.c_int64


def build(force=False):
    src = os.path.join(HERE, "ndsm_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-s", "-C", HERE] + (["-B"] if force else []), check=True, env=env)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.orc_relax3d.argtypes = [c.c_char_p, i64, i64, i64, vp, vp, vp, vp, vp]
        L.orc_residual3d.argtypes = [c.c_char_p, i64, i64, i64, vp, vp, vp, vp, vp, vp]
        L.orc_relax_nd.argtypes = [c.c_int, vp, vp, c.c_char_p, vp, vp]
        L.orc_residual_nd.argtypes = [c.c_int, vp, vp, c.c_char_p, vp, vp, vp]
        L.orc_mg_interp.argtypes = [c.c_int, vp, vp, vp, vp, vp, vp]
        L.orc_mg_restrict.argtypes = [c.c_int, vp, vp, vp, vp, vp, vp]
        L.orc_update_u.argtypes = [i64, vp, vp, vp, vp]
        L.orc_du_metrics.argtypes = [i64, vp, vp, vp]
        L.orc_mean.argtypes = [i64, vp]
        L.orc_mean.restype = c.c_double
        L.orc_ngrids.argtypes = [i64]
        L.orc_mg_new.argtypes = [c.c_int, vp, c.c_int, vp, c.c_int, i64]
        L.orc_mg_new.restype = vp
        L.orc_mg_delete.argtypes = [vp]
        L.orc_mg_set.argtypes = [vp, i64, c.c_double, c.c_char_p]
        L.orc_mg_shape.argtypes = [vp, c.c_int, vp]
        L.orc_mg_mesh.argtypes = [vp, c.c_int, c.c_int]
        L.orc_mg_mesh.restype = c.POINTER(c.c_double)
        L.orc_mg_u.argtypes = [vp, c.c_int]
        L.orc_mg_u.restype = c.POINTER(c.c_double)
        L.orc_mg_rhs.argtypes = [vp, c.c_int]
        L.orc_mg_rhs.restype = c.POINTER(c.c_double)
        L.orc_mg_load.argtypes = [vp, vp, vp]
        L.orc_mg_relax.argtypes = [vp, c.c_int]
        L.orc_mg_residual.argtypes = [vp, c.c_int, vp]
        L.orc_mg_solve_exact.argtypes = [vp, c.c_int]
        L.orc_mg_last_nexact.argtypes = [vp]
        L.orc_v_cycle.argtypes = [vp]
        L.orc_solve_poisson_bvp.argtypes = [vp, c.c_double, i64, vp, vp, vp, vp]
        L.orc_trace_du.argtypes = [c.c_int, c.c_int]
        L.orc_trace_du.restype = c.c_double
        L.orc_trace_ncycles.argtypes = [c.c_int]
        L.orc_trace_nexact.argtypes = [c.c_int, c.c_int]
        L.orc_trapz2d.argtypes = [i64, i64, c.c_double, c.c_double, vp]
        L.orc_trapz2d.restype = c.c_double
        L.orc_compute_At.argtypes = [vp, vp, c.c_double, c.c_int, vp, vp]
        L.orc_add_flux_balance_fields.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.orc_curl.argtypes = [vp, vp, vp, vp]
        L.orc_bc_setup.argtypes = [vp] * 11
        L.orc_interp_table.argtypes = [i64, vp, i64, vp, vp, vp, vp]
        L.orc_restrict_table.argtypes = [i64, vp, i64, vp, vp, vp, vp, vp]
        L.orc_bracket.argtypes = [vp, i64, c.c_double, vp, vp, vp]
        L.orc_set_num_threads.argtypes = [c.c_int]
        L.orc_set_num_threads.restype = None
        L.ndsm_vector_solve.argtypes = [c.c_size_t] + [vp] * 8
        L.orc_poisson_solve.argtypes = [c.c_int, vp, c.c_char_p, i64, i64, i64, c.c_int, c.c_double, c.c_double,
                                        vp, vp, vp, vp, vp, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _meshptrs(mesh):
    arr = (vp * len(mesh))(*[m.ctypes.data for m in mesh])
    return arr


def _i64(v):
    return np.array(v, dtype=np.int64)


# ---- single operators (dense numpy arrays in numpy order (nz,ny,nx) / (ny,nx)) -------------------

def relax3d(copt, mesh, rhs, u, nsweeps=1):
    x, y, z = [_f(m) for m in mesh]
    u = _f(u).copy()
    rhs = _f(rhs)
    for _ in range(nsweeps):
        lib().orc_relax3d(copt.encode(), x.size, y.size, z.size, _p(x), _p(y), _p(z), _p(rhs), _p(u))
    return u


def residual3d(copt, mesh, rhs, u):
    x, y, z = [_f(m) for m in mesh]
    u = _f(u)
    rhs = _f(rhs)
    r = np.zeros_like(u)
    lib().orc_residual3d(copt.encode(), x.size, y.size, z.size, _p(x), _p(y), _p(z), _p(rhs), _p(u), _p(r))
    return r


def relax_nd(copt, mesh, rhs, u, nsweeps=1):
    mesh = [_f(m) for m in mesh]
    nshape = _i64([m.size for m in mesh])
    u = _f(u).copy()
    rhs = _f(rhs)
    for _ in range(nsweeps):
        lib().orc_relax_nd(len(mesh), _p(nshape), _meshptrs(mesh), copt.encode(), _p(u), _p(rhs))
    return u


def residual_nd(copt, mesh, rhs, u):
    mesh = [_f(m) for m in mesh]
    nshape = _i64([m.size for m in mesh])
    u = _f(u)
    rhs = _f(rhs)
    r = np.zeros_like(u)
    lib().orc_residual_nd(len(mesh), _p(nshape), _meshptrs(mesh), copt.encode(), _p(u), _p(rhs), _p(r))
    return r


def mg_interp(mesh_f, mesh_c, u_c, mode=1):
    mesh_f = [_f(m) for m in mesh_f]
    mesh_c = [_f(m) for m in mesh_c]
    nf = _i64([m.size for m in mesh_f])
    nc = _i64([m.size for m in mesh_c])
    u_c = _f(u_c)
    u_f = np.zeros(tuple(int(v) for v in nf[::-1]))
    lib().orc_set_transfer_mode(mode)
    lib().orc_mg_interp(len(mesh_f), _p(nf), _meshptrs(mesh_f), _p(nc), _meshptrs(mesh_c), _p(u_c), _p(u_f))
    lib().orc_set_transfer_mode(1)
    return u_f


def mg_restrict(mesh_f, mesh_c, u_f, mode=1):
    mesh_f = [_f(m) for m in mesh_f]
    mesh_c = [_f(m) for m in mesh_c]
    nf = _i64([m.size for m in mesh_f])
    nc = _i64([m.size for m in mesh_c])
    u_f = _f(u_f)
    u_c = np.zeros(tuple(int(v) for v in nc[::-1]))
    lib().orc_set_transfer_mode(mode)
    lib().orc_mg_restrict(len(mesh_f), _p(nf), _meshptrs(mesh_f), _p(nc), _meshptrs(mesh_c), _p(u_f), _p(u_c))
    lib().orc_set_transfer_mode(1)
    return u_c


def update_u(u_old, u_new):
    a = _f(u_new).copy()
    b = _f(u_old)
    dmax = c.c_double(0)
    dmean = c.c_double(0)
    lib().orc_update_u(a.size, _p(b), _p(a), c.byref(dmax), c.byref(dmean))
    return a, dmax.value, dmean.value


class OracleMG:
    """MG_HANDLE of the oracle (orc_mg)."""

    def __init__(self, mesh, copt, ms=5, ex_tol=1e-13, du_max=True, nmax_exact=10000, ngrids=-1):
        self.L = lib()
        self.mesh = [_f(m) for m in mesh]
        self.ndim = len(self.mesh)
        nshape = _i64([m.size for m in self.mesh])
        if ngrids < 0:
            ngrids = self.L.orc_ngrids(int(nshape.min()))
        self.h = self.L.orc_mg_new(self.ndim, _p(nshape), ngrids, _meshptrs(self.mesh), int(bool(du_max)),
                                   int(nmax_exact))
        if not self.h:
            raise ValueError("orc_mg_new failed")
        self.ngrids = ngrids
        self.L.orc_mg_set(self.h, ms, ex_tol, copt.encode())

    def close(self):
        if getattr(self, "h", None):
            self.L.orc_mg_delete(self.h)
            self.h = None

    __del__ = close

    def shape(self, g):
        s = np.zeros(3, dtype=np.int64)
        self.L.orc_mg_shape(self.h, g, _p(s))
        return tuple(int(v) for v in s[: self.ndim])

    def level_mesh(self, g):
        return [np.ctypeslib.as_array(self.L.orc_mg_mesh(self.h, g, d), shape=(n,)).copy()
                for d, n in enumerate(self.shape(g))]

    def load(self, u, rhs):
        u = _f(u)
        rhs = _f(rhs)
        self.L.orc_mg_load(self.h, _p(u), _p(rhs))

    def u(self, g):
        shp = self.shape(g)[::-1]
        return np.ctypeslib.as_array(self.L.orc_mg_u(self.h, g), shape=shp).copy()

    def rhs(self, g):
        shp = self.shape(g)[::-1]
        return np.ctypeslib.as_array(self.L.orc_mg_rhs(self.h, g), shape=shp).copy()

    def v_cycle(self):
        self.L.orc_v_cycle(self.h)
        return self.L.orc_mg_last_nexact(self.h)

    def solve_exact(self, g):
        self.L.orc_mg_solve_exact(self.h, g)
        return self.L.orc_mg_last_nexact(self.h)

    def solve(self, u, rhs=None, vc_tol=1e-10, nmax=1024):
        u = _f(u).copy()
        rhs = np.zeros_like(u) if rhs is None else _f(rhs)
        du = c.c_double(0)
        ierr = c.c_int64(0)
        self.L.orc_trace_reset()
        self.L.orc_solve_poisson_bvp(self.h, vc_tol, nmax, _p(u), _p(rhs), c.byref(du), c.byref(ierr))
        return int(ierr.value), u, du.value, self.L.orc_trace_ncycles(0)


# ---- driver -------------------------------------------------------------------------------------

SOLVE_NAMES = ("chi1", "chi2", "chi3", "chi4", "chi5", "chi6", "Ax", "Ay", "Az")


def vector_potential(x, y, z, b, niterex_max=10000, ncycles_max=1024, ex_tol=1e-13, vc_tol=1e-10, ms=5, mean=False,
                     debug=False, flxcrl=0, A0=None, trace=False):
    """Same call as ndsm.vector_potential (ndsm.py:66) against the oracle library."""
    L = lib()
    b = _f(b)
    x, y, z = _f(x), _f(y), _f(z)
    nshape = np.array(b.shape[::-1], dtype=np.intc)
    ioptc = np.zeros(16, dtype=np.intc)
    ropt = np.zeros(16)
    ioptc[L.get_iopt_ms()] = ms
    ioptc[L.get_iopt_ncycles()] = ncycles_max
    ioptc[L.get_iopt_iopt_nmaxex()] = niterex_max
    ropt[L.get_ropt_vtol()] = vc_tol
    ropt[L.get_ropt_ctol()] = ex_tol
    ioptc[L.get_iopt_debug()] = 1 if debug else 0
    ioptc[L.get_iopt_dumax()] = 0 if mean else 1
    ioptc[4] = flxcrl
    A = np.zeros(b.size) if A0 is None else _f(A0).ravel().copy()
    bf = b.flatten()
    L.orc_trace_reset()
    ierr = L.ndsm_vector_solve(c.c_size_t(b.size), _p(nshape), _p(ioptc), _p(ropt), _p(x), _p(y), _p(z), _p(A), _p(bf))
    shp = tuple(int(v) for v in nshape[::-1])
    res = (ierr, A.reshape(shp), bf.reshape(shp))
    if trace:
        info = {}
        for s, name in enumerate(SOLVE_NAMES):
            nc = L.orc_trace_ncycles(s)
            info[name] = {"du": [L.orc_trace_du(s, k) for k in range(nc)],
                          "nexact": [L.orc_trace_nexact(s, k) for k in range(nc)]}
        info["seconds"] = float(ropt[2])
        return res + (info,)
    return res


def bc_setup(x, y, z, b, niterex_max=10000, ncycles_max=1024, ex_tol=1e-13, vc_tol=1e-10, ms=5, mean=False):
    """BC-setup stage only: returns (phi[6], chi[6], At1[6], At2[6]) with faces as (n2, n1) numpy arrays."""
    L = lib()
    b = _f(b)
    x, y, z = _f(x), _f(y), _f(z)
    nshape = np.array(b.shape[::-1], dtype=np.intc)
    ioptc = np.zeros(16, dtype=np.intc)
    ropt = np.zeros(16)
    ioptc[0], ioptc[1], ioptc[7], ioptc[6] = ms, ncycles_max, niterex_max, (0 if mean else 1)
    ropt[0], ropt[1] = vc_tol, ex_tol
    nc = [(1, 2), (1, 2), (0, 2), (0, 2), (0, 1), (0, 1)]
    faces = [[np.zeros((int(nshape[b2]), int(nshape[a]))) for (a, b2) in nc] for _ in range(3)]
    arrs = [(vp * 6)(*[f.ctypes.data for f in faces[q]]) for q in range(3)]
    phi = np.zeros(6)
    L.orc_trace_reset()
    L.orc_bc_setup(_p(nshape), _p(ioptc), _p(ropt), _p(x), _p(y), _p(z), _p(b), _p(phi), arrs[0], arrs[1], arrs[2])
    return phi, faces[0], faces[1], faces[2]


def flux_curl(x, y, z, phi, A, flxcrl=0):
    """add_flux_balance_fields + curl in the reference order.  Returns (A, B)."""
    L = lib()
    x, y, z = _f(x), _f(y), _f(z)
    A = _f(A).copy()
    B = np.zeros_like(A)
    nshape = _i64([x.size, y.size, z.size])
    dq = np.array([x[1] - x[0], y[1] - y[0], z[1] - z[0]])
    phi = _f(phi)
    if flxcrl == 1:
        L.orc_curl(_p(nshape), _p(dq), _p(A), _p(B))
        L.orc_add_flux_balance_fields(_p(nshape), _p(x), _p(y), _p(z), _p(phi), _p(B), _p(A))
    else:
        L.orc_add_flux_balance_fields(_p(nshape), _p(x), _p(y), _p(z), _p(phi), _p(B), _p(A))
        L.orc_curl(_p(nshape), _p(dq), _p(A), _p(B))
    return A, B


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def poisson_solve(mesh, copt, u, rhs=None, ms=5, ncycles_max=1024, niterex_max=10000, mean=False, vc_tol=1e-10,
                  ex_tol=1e-13):
    """solve_poisson_bvp (ndsm_poisson.f90:63) on a caller-defined 2D/3D problem; u, rhs in numpy order
    (nz,ny,nx).  Returns (ierr, u, du_last, ncycles)."""
    mesh = [_f(m) for m in mesh]
    nshape = np.array([m.size for m in mesh], dtype=np.intc)
    u = np.array(u, dtype=np.float64, order="C")
    r = np.zeros_like(u) if rhs is None else _f(rhs)
    du, nc = c.c_double(0), c.c_int(0)
    z = mesh[2] if len(mesh) > 2 else None
    ierr = lib().orc_poisson_solve(len(mesh), _p(nshape), copt.encode(), ms, ncycles_max, niterex_max,
                                   0 if mean else 1, vc_tol, ex_tol, _p(mesh[0]), _p(mesh[1]), _p(z), _p(u), _p(r),
                                   c.byref(du), c.byref(nc))
    return ierr, u, du.value, nc.value
