"""Times ndsm_b200_vector_solve_rank (513^3 dipole by default, faces resident in HBM) under several scheduling
configurations of the three component solves in ONE launch of the ranks -- environment variables of the library
are read per call, so every configuration rebuilds its hierarchies and graphs in the warm-up calls.

torchrun --nproc-per-node N scripts/sweep_groups.py [--n 513] [--steps 3] [--warmup 2] [--configs "a;b;c"]
A configuration is a comma-free list of KEY=VALUE pairs separated by spaces ("" = defaults).
Prints one line per configuration on rank 0: max-over-ranks wall ms per solve (barrier on both sides), the
library's own stage times, V-cycle counts, and whether every rank's slab equals the first configuration's bit for bit.
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

DEFAULT = ";".join([
    "",
    "NDSM_COMPONENT_GROUPS=012",
    "NDSM_COMPONENT_GROUPS=01|2",
    "NDSM_B200_ZCHUNK_FORCE=16",
])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=513)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--configs", default=DEFAULT)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from ndsm_b200 import load_library, synthetic
    from ndsm_b200 import dist as ndist
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["NDSM_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load_library()
    ndist.init_from_torch(local)
    n = args.n
    x, y, z = synthetic.mesh(n, n, n)
    b = synthetic.dipole(x, y, z)
    k0, k1 = ndist.slab_range(n, world, rank)
    faces_d = [torch.from_numpy(f).cuda() for f in ndist.extract_faces(b)]
    fptr = [f.data_ptr() for f in faces_d]
    dA = torch.empty((3, k1 - k0, n, n), dtype=torch.float64, device="cuda")
    dB = torch.empty_like(dA)
    ref = None
    for cfg in args.configs.split(";"):
        pairs = [kv.split("=", 1) for kv in cfg.replace("|", ",").split() if "=" in kv]
        saved = {k: os.environ.get(k) for k, _ in pairs}
        for k, v in pairs:
            os.environ[k] = v
        try:
            def step():
                rc, _, _, _ = ndist.vector_potential_rank(x, y, z, fptr, out=(dA.data_ptr(), dB.data_ptr()),
                                                          faces_on_device=True)
                if rc != 0:
                    raise RuntimeError("rc=%d" % rc)
            for _ in range(args.warmup):
                step()
            times = []
            stage = np.zeros(8)
            for _ in range(args.steps):
                torch.cuda.synchronize(); dist.barrier()
                t0 = time.perf_counter()
                step()
                torch.cuda.synchronize()
                t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times.append(float(t.item()))
                tim = np.zeros(8)
                lib.ndsm_b200_last_timing(tim.ctypes.data_as(ctypes.c_void_p))
                stage += tim
            stage /= args.steps
            mode = lib.ndsm_b200_last_components_mode()
            cyc = [lib.ndsm_b200_trace_ncycles(6 + c) for c in range(3)]
            if ref is None:
                ref = (dA.clone(), dB.clone())
                same = True
            else:
                same = bool(torch.equal(dA, ref[0]) and torch.equal(dB, ref[1]))
            ok = torch.tensor([1 if same else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if rank == 0:
                print("SWEEP world=%d n=%d cfg=[%s] mode=%d ms=%s min=%.2f bc=%.2f solve3d=%.2f post=%.2f cycles=%s slabs_equal_first=%s"
                      % (world, n, cfg, mode, ["%.2f" % v for v in times], min(times), stage[2], stage[3], stage[4], cyc,
                         bool(ok.item())), flush=True)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    dist.barrier()
    ndist.shutdown() if hasattr(ndist, "shutdown") else None
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
