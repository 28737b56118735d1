"""z-slab domain decomposition checked on ONE GPU: the solve runs with G virtual ranks on the same device
(NDSM_VIRTUAL_SLABS, same kernels and halo logic as the multi-process NCCL path, device copies instead of
NVLink traffic) and must reproduce the single-slab result -- bit for bit with the max metric."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture
def slab_env():
    saved = {k: os.environ.get(k) for k in ("NDSM_VIRTUAL_SLABS", "NDSM_SLAB_MIN_PLANES")}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def solve(shape, world, min_planes, **kw):
    from ndsm_b200 import synthetic, vector_potential
    os.environ["NDSM_VIRTUAL_SLABS"] = str(world)
    os.environ["NDSM_SLAB_MIN_PLANES"] = str(min_planes)
    x, y, z = synthetic.mesh(*shape)
    b = synthetic.dipole(x, y, z)
    return vector_potential(x, y, z, b, trace=True, **kw)


@pytest.mark.parametrize("shape", [(44, 44, 44), (40, 33, 52), (65, 65, 65)])
@pytest.mark.parametrize("world,min_planes", [(2, 16), (2, 4), (3, 4), (4, 4)])
def test_virtual_slabs_reproduce_single_slab(gpu_lib, slab_env, shape, world, min_planes):
    ref = solve(shape, 1, 16)
    got = solve(shape, world, min_planes)
    assert got[0] == ref[0] == 0
    for name in ("Ax", "Ay", "Az"):
        assert got[3][name]["du"] == ref[3][name]["du"], name      # identical du history (max metric)
        assert got[3][name]["nexact"] == ref[3][name]["nexact"], name
    assert np.array_equal(got[1], ref[1])
    assert np.array_equal(got[2], ref[2])


def test_virtual_slabs_mean_metric_and_flxcrl(gpu_lib, slab_env):
    ref = solve((44, 40, 48), 1, 16, mean=True, flxcrl=1)
    got = solve((44, 40, 48), 3, 4, mean=True, flxcrl=1)
    for name in ("Ax", "Ay", "Az"):
        assert abs(len(got[3][name]["du"]) - len(ref[3][name]["du"])) <= 1
    assert rel_err(got[1], ref[1]) <= 1e-12
    assert rel_err(got[2], ref[2]) <= 1e-12


def test_virtual_slabs_with_initial_guess(gpu_lib, slab_env):
    from ndsm_b200 import synthetic, vector_potential
    x, y, z = synthetic.mesh(40, 36, 44)
    b = synthetic.dipole(x, y, z)
    A0 = 0.01 * np.random.default_rng(7).standard_normal(b.shape)
    os.environ["NDSM_VIRTUAL_SLABS"] = "1"
    ref = vector_potential(x, y, z, b, A0=A0)
    os.environ["NDSM_VIRTUAL_SLABS"] = "2"
    os.environ["NDSM_SLAB_MIN_PLANES"] = "4"
    got = vector_potential(x, y, z, b, A0=A0)
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
