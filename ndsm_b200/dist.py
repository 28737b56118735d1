"""Multi-GPU host side: one process per GPU, z-slab decomposition (include/ndsm_b200.h, "Multi-GPU").

The library brings its own NCCL communicator; torch.distributed (or anything else) is only used to hand the
128-byte NCCL unique id from rank 0 to the other ranks.
"""
import ctypes

import numpy as np

from .lib_loader import load_library
from .ndsm import _options


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def init_from_torch(device=None):
    """Create the library's communicator inside an initialised torch.distributed job.  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    lib = load_library()
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        rc = lib.ndsm_b200_dist_unique_id(_ptr(uid))
        if rc != 0:
            raise RuntimeError("ndsm_b200_dist_unique_id failed (%d): NCCL not available" % rc)
    if dist.get_backend() == "nccl":
        t = torch.from_numpy(uid).cuda(device)
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()
    else:
        t = torch.from_numpy(uid)
        dist.broadcast(t, 0)
        uid = t.numpy()
    uid = np.ascontiguousarray(uid)
    rc = lib.ndsm_b200_dist_init(rank, world, _ptr(uid))
    if rc != 0:
        raise RuntimeError("ndsm_b200_dist_init failed with code %d" % rc)
    return rank, world


def slab_range(nz, world, rank):
    lib = load_library()
    k0, k1 = ctypes.c_int(0), ctypes.c_int(0)
    assert lib.ndsm_b200_slab_range(int(nz), int(world), int(rank), ctypes.byref(k0), ctypes.byref(k1)) == 0
    return k0.value, k1.value


def extract_faces(b):
    """Six boundary-normal faces of b (3,nz,ny,nx) in the library's dense layout (extract_bn,
    ndsm_vector_potential.f90:283-293): x-faces (nz,ny), y-faces (nz,nx), z-faces (ny,nx) in numpy order."""
    return [np.ascontiguousarray(a) for a in (b[0][:, :, 0], b[0][:, :, -1], b[1][:, 0, :], b[1][:, -1, :],
                                              b[2][0, :, :], b[2][-1, :, :])]


def vector_potential_rank(x, y, z, faces, niterex_max=10000, ncycles_max=1024, ex_tol=1e-13, vc_tol=1e-10, ms=5,
                          mean=False, debug=False, flxcrl=0, out=None, faces_on_device=False):
    """This rank's z-slab of A and B.  `faces`: six arrays (numpy, or device pointers when faces_on_device).
    `out`: optional (A_ptr, B_ptr) device pointers; otherwise numpy slabs (3, k1-k0, ny, nx) are returned.
    Returns (ierr, A_slab, B_slab, (k0, k1))."""
    lib = load_library()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    nshape = np.array([x.size, y.size, z.size, 3], dtype=np.intc)
    ioptc, ropt = _options(lib, niterex_max, ncycles_max, ex_tol, vc_tol, ms, mean, debug, flxcrl)
    world, rank = lib.ndsm_b200_dist_world(), lib.ndsm_b200_dist_rank()
    k0, k1 = slab_range(z.size, world, rank)
    if faces_on_device:
        fp = (ctypes.c_void_p * 6)(*[ctypes.c_void_p(int(p)) for p in faces])
    else:
        faces = [np.ascontiguousarray(f, dtype=np.float64) for f in faces]
        fp = (ctypes.c_void_p * 6)(*[f.ctypes.data for f in faces])
    if out is None:
        A = np.zeros((3, k1 - k0, y.size, x.size))
        B = np.zeros_like(A)
        rc = lib.ndsm_b200_vector_solve_rank(_ptr(nshape), _ptr(ioptc), _ptr(ropt), _ptr(x), _ptr(y), _ptr(z), fp,
                                             int(faces_on_device), _ptr(A), _ptr(B), 0)
    else:
        A, B = out
        rc = lib.ndsm_b200_vector_solve_rank(_ptr(nshape), _ptr(ioptc), _ptr(ropt), _ptr(x), _ptr(y), _ptr(z), fp,
                                             int(faces_on_device), ctypes.c_void_p(int(A)), ctypes.c_void_p(int(B)), 1)
    return rc, A, B, (k0, k1)


def poisson_solve(x, y, z, u, rhs=None, copt="NDDNDD", ms=5, ncycles_max=1024, niterex_max=10000, mean=False,
                  vc_tol=1e-10, ex_tol=1e-13):
    """solve_poisson_bvp (ndsm_poisson.f90:63) on full host arrays u, rhs of shape (nz,ny,nx) through
    ndsm_b200_poisson_solve (single GPU; with NDSM_VIRTUAL_SLABS=G the z-slab path with G virtual ranks).
    Returns (ierr, u, du_last, ncycles)."""
    lib = load_library()
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    u = np.array(u, dtype=np.float64, order="C")
    nshape = np.array([x.size, y.size, z.size], dtype=np.intc)
    r = None if rhs is None else np.ascontiguousarray(rhs, dtype=np.float64)
    du, nc = ctypes.c_double(0), ctypes.c_int(0)
    ierr = lib.ndsm_b200_poisson_solve(3, _ptr(nshape), copt.encode(), ms, ncycles_max, niterex_max, 0 if mean else 1,
                                       vc_tol, ex_tol, _ptr(x), _ptr(y), _ptr(z), _ptr(u),
                                       None if r is None else _ptr(r), ctypes.byref(du), ctypes.byref(nc))
    return ierr, u, du.value, nc.value


def poisson_solve_rank(x, y, z, u_slab_ptr, rhs_slab_ptr=None, copt="NDDNDD", ms=5, ncycles_max=1024,
                       niterex_max=10000, mean=False, vc_tol=1e-10, ex_tol=1e-13):
    """This rank's z-slab of a 3D scalar Poisson solve (ndsm_b200_poisson_solve_rank).  u_slab_ptr / rhs_slab_ptr
    are DEVICE pointers to the dense planes [k0,k1) = slab_range(nz, world, rank) (u in/out, rhs may be None).
    Returns (ierr, du_last, ncycles)."""
    lib = load_library()
    x, y, z = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z))
    nshape = np.array([x.size, y.size, z.size], dtype=np.intc)
    du, nc = ctypes.c_double(0), ctypes.c_int(0)
    ierr = lib.ndsm_b200_poisson_solve_rank(_ptr(nshape), copt.encode(), ms, ncycles_max, niterex_max,
                                            0 if mean else 1, vc_tol, ex_tol, _ptr(x), _ptr(y), _ptr(z),
                                            ctypes.c_void_p(int(u_slab_ptr)),
                                            None if rhs_slab_ptr is None else ctypes.c_void_p(int(rhs_slab_ptr)),
                                            ctypes.byref(du), ctypes.byref(nc))
    return ierr, du.value, nc.value
