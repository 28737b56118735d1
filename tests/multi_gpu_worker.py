"""torchrun worker: every rank solves its z-slab (peer-memory transport over NVLink by default, NCCL send/recv with
NDSM_P2P=0) and compares it with a full single-GPU solve of the same problem done locally through the frozen ABI.
Prints 'MULTI_GPU_OK <rank>' on success."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ndsm_b200 import synthetic, vector_potential  # noqa: E402
from ndsm_b200 import dist as ndist  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["NDSM_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["NDSM_SLAB_MIN_POINTS"] = "0"  # partition these small grids too (production keeps them replicated)
    rank, world = ndist.init_from_torch(local)
    from ndsm_b200 import load_library
    lib = load_library()
    shapes = [(65, 65, 65), (72, 40, 96)]
    for shape in shapes:
        for mean in (False, True):
            x, y, z = synthetic.mesh(*shape)
            b = synthetic.dipole(x, y, z)
            ierr, A, B, (k0, k1) = ndist.vector_potential_rank(x, y, z, ndist.extract_faces(b), mean=mean)
            nd = lib.ndsm_b200_last_partitioned_levels()
            rerr, Ar, Br = vector_potential(x, y, z, b, mean=mean)
            assert ierr == rerr == 0, (ierr, rerr)
            assert nd > 0 or world == 1, nd   # the finest levels really are partitioned over the ranks
            assert A.shape == (3, k1 - k0, shape[1], shape[0])
            if mean:
                ra = np.abs(A - Ar[:, k0:k1]).max() / np.abs(Ar).max()
                rb = np.abs(B - Br[:, k0:k1]).max() / np.abs(Br).max()
                assert ra <= 1e-12 and rb <= 1e-12, (ra, rb)
            else:  # max metric: the slab path is bit-identical to the single-GPU path
                assert np.array_equal(A, Ar[:, k0:k1]), np.abs(A - Ar[:, k0:k1]).max()
                assert np.array_equal(B, Br[:, k0:k1]), np.abs(B - Br[:, k0:k1]).max()
    # scalar Poisson solve (BASELINE config 5) on z-slabs over the world communicator vs the single-GPU entry
    nx, ny, nz = 48, 40, 24 * world
    x = np.linspace(0, 1, nx)
    dx = x[1] - x[0]
    y, z = np.arange(ny) * dx, np.arange(nz) * dx
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    uex = np.cos(np.pi * X) * np.sin(np.pi * Y / y[-1]) * np.sin(np.pi * Z / z[-1])
    rhs = -(np.pi ** 2) * (1 + 1 / y[-1] ** 2 + 1 / z[-1] ** 2) * uex
    k0, k1 = ndist.slab_range(nz, world, rank)
    du_s = torch.zeros((k1 - k0, ny, nx), dtype=torch.float64, device="cuda")
    dr_s = torch.from_numpy(np.ascontiguousarray(rhs[k0:k1])).cuda()
    ierr, du, nc = ndist.poisson_solve_rank(x, y, z, du_s.data_ptr(), dr_s.data_ptr())
    assert lib.ndsm_b200_last_partitioned_levels() > 0
    rerr, ur, rdu, rnc = ndist.poisson_solve(x, y, z, np.zeros_like(uex), rhs)
    assert ierr == rerr == 0 and nc == rnc and du == rdu, (ierr, rerr, nc, rnc, du, rdu)
    assert np.array_equal(du_s.cpu().numpy(), ur[k0:k1])
    dist.barrier()
    print("MULTI_GPU_OK %d of %d (%s)" % (rank, world, lib.ndsm_b200_dist_transport().decode()), flush=True)
    from ndsm_b200 import load_library
    load_library().ndsm_b200_dist_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
