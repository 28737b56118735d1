"""diagnostic: run-to-run determinism of the single-slab solve and slab-vs-single identity under env toggles"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndsm_b200 import synthetic, vector_potential, load_library
n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
x, y, z = synthetic.mesh(n)
b = synthetic.dipole(x, y, z)
os.environ["NDSM_SLAB_MIN_POINTS"] = "0"
def run(**env):
    for k, v in env.items():
        os.environ[k] = str(v)
    r = vector_potential(x, y, z, b, trace=True)
    for k in env:
        os.environ.pop(k, None)
    return r
ref = run()
ref2 = run()
print("single-slab run-to-run: A equal", np.array_equal(ref[1], ref2[1]), "du equal", all(ref[3][k]["du"] == ref2[3][k]["du"] for k in ("Ax", "Ay", "Az")),
      "chi du equal", all(ref[3]["chi%d" % f]["du"] == ref2[3]["chi%d" % f]["du"] for f in range(1, 7)))
for env in [dict(NDSM_VIRTUAL_SLABS=2), dict(NDSM_VIRTUAL_SLABS=2, NDSM_HALO_ONE_COLOUR=0), dict(NDSM_VIRTUAL_SLABS=2, NDSM_HALO_EXACT_DEPTH=0),
            dict(NDSM_VIRTUAL_SLABS=2, NDSM_HALO_ONE_COLOUR=0, NDSM_HALO_EXACT_DEPTH=0), dict(NDSM_VIRTUAL_SLABS=3), dict(NDSM_B200_SMALL=0)]:
    got = run(**env)
    nd = load_library().ndsm_b200_last_partitioned_levels()
    print(env, "nd", nd, "A equal", np.array_equal(got[1], ref[1]), "maxdiff %.3e" % np.abs(got[1] - ref[1]).max(),
          "du0", [got[3][k]["du"][0] == ref[3][k]["du"][0] for k in ("Ax", "Ay", "Az")], flush=True)
