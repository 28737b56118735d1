"""ctypes view of the MG_HANDLE operator seam and of the host-only planner (include/ndsm_b200.h §3)."""
import ctypes

import numpy as np

from .lib_loader import load_library


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class MGHandle:
    """GPU multigrid handle (new_mg_handle, ndsm_multigrid_core.f90:165).  Arrays are dense numpy arrays in
    numpy order (nz, ny, nx) / (ny, nx), i.e. the Fortran (nx, ny[, nz]) layout of the reference."""

    U, RHS, R = 0, 1, 2

    def __init__(self, mesh, copt, ms=5, ex_tol=1e-13, du_max=True, nmax_exact=10000, ngrids=-1):
        self.lib = load_library()
        self.mesh = [_f(m) for m in mesh]
        self.ndim = len(self.mesh)
        nshape = np.array([m.size for m in self.mesh], dtype=np.intc)
        m3 = self.mesh + [None] * (3 - self.ndim)
        self.h = self.lib.ndsm_b200_new_mg_handle(self.ndim, _ptr(nshape), ngrids, _ptr(m3[0]), _ptr(m3[1]),
                                                  _ptr(m3[2]), int(bool(du_max)), int(nmax_exact))
        if not self.h:
            raise RuntimeError("ndsm_b200_new_mg_handle failed (no CUDA device or invalid shape)")
        self.set_options(ms, ex_tol, copt)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ndsm_b200_delete_mg_handle(self.h)
            self.h = None

    __del__ = close

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed with code %d" % (what, rc))

    def set_options(self, ms, ex_tol, copt):
        self._chk(self.lib.ndsm_b200_mg_set_options(self.h, int(ms), float(ex_tol), copt.encode()), "set_options")

    @property
    def ngrids(self):
        return self.lib.ndsm_b200_mg_ngrids(self.h)

    def shape(self, level):
        s = np.zeros(3, dtype=np.intc)
        self._chk(self.lib.ndsm_b200_mg_level_shape(self.h, level, _ptr(s)), "level_shape")
        return tuple(int(v) for v in s[: self.ndim])  # (nx, ny[, nz])

    def level_mesh(self, level):
        out = []
        for d, n in enumerate(self.shape(level)):
            m = np.zeros(n)
            self._chk(self.lib.ndsm_b200_mg_level_mesh(self.h, level, d, _ptr(m)), "level_mesh")
            out.append(m)
        return out

    def put(self, which, level, arr):
        a = _f(arr)
        assert a.shape == self.shape(level)[::-1], (a.shape, self.shape(level))
        self._chk(self.lib.ndsm_b200_mg_put(self.h, which, level, _ptr(a)), "put")

    def get(self, which, level):
        a = np.zeros(self.shape(level)[::-1])
        self._chk(self.lib.ndsm_b200_mg_get(self.h, which, level, _ptr(a)), "get")
        return a

    def relax(self, level, nsweeps=1):
        self._chk(self.lib.ndsm_b200_mg_relax(self.h, level, nsweeps), "relax")

    def residual(self, level):
        self._chk(self.lib.ndsm_b200_mg_residual(self.h, level), "residual")
        return self.get(self.R, level)

    def restrict(self, level):
        self._chk(self.lib.ndsm_b200_mg_restrict(self.h, level), "restrict")
        return self.get(self.RHS, level + 1)

    def residual_restrict(self, level):
        """rhs(level+1) = R (rhs - L u) in one step; returns (rhs_coarse, fused?)."""
        f = ctypes.c_int(0)
        self._chk(self.lib.ndsm_b200_mg_residual_restrict(self.h, level, ctypes.byref(f)), "residual_restrict")
        return self.get(self.RHS, level + 1), bool(f.value)

    def interp_add(self, level):
        self._chk(self.lib.ndsm_b200_mg_interp_add(self.h, level), "interp_add")
        return self.get(self.U, level - 1)

    def solve_exact(self, level):
        it = ctypes.c_int(0)
        self._chk(self.lib.ndsm_b200_mg_solve_exact(self.h, level, ctypes.byref(it)), "solve_exact")
        return it.value

    def v_cycle(self):
        self._chk(self.lib.ndsm_b200_mg_v_cycle(self.h), "v_cycle")

    def solve(self, u, rhs=None, vc_tol=1e-10, nmax=1024):
        """solve_poisson_bvp.  Returns (ierr, u, du_last, ncycles)."""
        u = _f(u).copy()
        r = None if rhs is None else _f(rhs)
        du = ctypes.c_double(0)
        nc = ctypes.c_int(0)
        ierr = self.lib.ndsm_b200_mg_solve(self.h, float(vc_tol), int(nmax), _ptr(u), _ptr(r), ctypes.byref(du),
                                           ctypes.byref(nc))
        if ierr not in (0, 1):
            raise RuntimeError("ndsm_b200_mg_solve failed with code %d" % ierr)
        return ierr, u, du.value, nc.value

    def update_u(self, u_old, u_new):
        """update_u (ndsm_multigrid_core.f90:1077): returns (u_new := u_old, du_max, du_mean)."""
        a = _f(u_new).copy()
        b = _f(u_old)
        dmax = ctypes.c_double(0)
        dmean = ctypes.c_double(0)
        self._chk(self.lib.ndsm_b200_mg_update_u(self.h, _ptr(b), _ptr(a), ctypes.byref(dmax), ctypes.byref(dmean)),
                  "update_u")
        return a, dmax.value, dmean.value


class Plan:
    """Host-only hierarchy/tables (no GPU needed)."""

    def __init__(self, mesh, ngrids=-1):
        self.lib = load_library()
        self.mesh = [_f(m) for m in mesh]
        self.ndim = len(self.mesh)
        nshape = np.array([m.size for m in self.mesh], dtype=np.intc)
        m3 = self.mesh + [None] * (3 - self.ndim)
        self.p = self.lib.ndsm_b200_plan_create(self.ndim, _ptr(nshape), ngrids, _ptr(m3[0]), _ptr(m3[1]), _ptr(m3[2]))
        if not self.p:
            raise ValueError("invalid shape for a multigrid hierarchy")

    def close(self):
        if getattr(self, "p", None):
            self.lib.ndsm_b200_plan_destroy(self.p)
            self.p = None

    __del__ = close

    @property
    def ngrids(self):
        return self.lib.ndsm_b200_plan_ngrids(self.p)

    def level(self, g):
        s = np.zeros(3, dtype=np.intc)
        lay = np.zeros(4, dtype=np.int64)
        w = np.zeros(5)
        assert self.lib.ndsm_b200_plan_level(self.p, g, _ptr(s), _ptr(lay), _ptr(w)) == 0
        return {"shape": tuple(int(v) for v in s), "hp": int(lay[0]), "mcnt": int(lay[1]), "ps": int(lay[2]),
                "cs": int(lay[3]), "wx": w[0], "wy": w[1], "wz": w[2], "w1": w[3], "wc": w[4]}

    def mesh_of(self, g, d):
        n = self.level(g)["shape"][d]
        m = np.zeros(n)
        assert self.lib.ndsm_b200_plan_mesh(self.p, g, d, _ptr(m)) == 0
        return m

    def interp_table(self, g, d):
        nf = self.level(g)["shape"][d]
        lo = np.zeros(nf, dtype=np.intc)
        wl = np.zeros(nf)
        wh = np.zeros(nf)
        assert self.lib.ndsm_b200_plan_interp(self.p, g, d, _ptr(lo), _ptr(wl), _ptr(wh)) == 0
        return lo, wl, wh

    def slab_partition(self, world, min_planes=16):
        """Returns (ndist, zs) with zs[level][rank] plane boundaries (ndist+1 rows when ndist > 0)."""
        nd = ctypes.c_int(0)
        zs = np.zeros((self.ngrids, world + 1), dtype=np.intc)
        assert self.lib.ndsm_b200_plan_slab_partition(self.p, world, min_planes, ctypes.byref(nd), _ptr(zs)) == 0
        n = nd.value
        return n, (zs[: n + 1].copy() if n > 0 else zs[:0])

    def restrict_table(self, g, d):
        nc = self.level(g + 1)["shape"][d]
        first = np.zeros(nc, dtype=np.intc)
        count = np.zeros(nc, dtype=np.intc)
        c2 = np.zeros((nc, 8))
        w2 = ctypes.c_double(0)
        assert self.lib.ndsm_b200_plan_restrict(self.p, g, d, _ptr(first), _ptr(count), _ptr(c2), ctypes.byref(w2)) == 0
        return first, count, c2, w2.value
