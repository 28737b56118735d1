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
    saved = {k: os.environ.get(k) for k in ("NDSM_VIRTUAL_SLABS", "NDSM_SLAB_MIN_PLANES", "NDSM_SLAB_MIN_POINTS",
                                            "NDSM_HALO_PLANES")}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def solve(shape, world, min_planes, halo=None, **kw):
    """Solve with `world` virtual ranks; the size threshold that keeps small levels replicated in production
    is switched off so that these small grids really are partitioned (asserted)."""
    from ndsm_b200 import load_library, synthetic, vector_potential
    os.environ["NDSM_VIRTUAL_SLABS"] = str(world)
    os.environ["NDSM_SLAB_MIN_PLANES"] = str(min_planes)
    os.environ["NDSM_SLAB_MIN_POINTS"] = "0"
    if halo is None:
        os.environ.pop("NDSM_HALO_PLANES", None)
    else:
        os.environ["NDSM_HALO_PLANES"] = str(halo)
    x, y, z = synthetic.mesh(*shape)
    b = synthetic.dipole(x, y, z)
    out = vector_potential(x, y, z, b, trace=True, **kw)
    nd = load_library().ndsm_b200_last_partitioned_levels()
    assert (nd > 0) == (world > 1), (world, nd)
    return out


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


@pytest.mark.parametrize("halo", [4, 5, 9, 12])
def test_halo_depth_does_not_change_the_result(gpu_lib, slab_env, halo):
    shape = (48, 40, 80)
    ref = solve(shape, 1, 16)
    got = solve(shape, 3, halo, halo=halo)
    for name in ("Ax", "Ay", "Az"):
        assert got[3][name]["du"] == ref[3][name]["du"], name
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
    os.environ["NDSM_SLAB_MIN_POINTS"] = "0"
    got = vector_potential(x, y, z, b, A0=A0)
    from ndsm_b200 import load_library
    assert load_library().ndsm_b200_last_partitioned_levels() > 0
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def poisson_case(shape):
    nx, ny, nz = shape
    x = np.linspace(0, 1, nx)
    dx = x[1] - x[0]
    y, z = np.arange(ny) * dx, np.arange(nz) * dx
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    uex = np.cos(np.pi * X) * np.sin(np.pi * Y / y[-1]) * np.sin(np.pi * Z / z[-1])
    rhs = -(np.pi ** 2) * (1 + 1 / y[-1] ** 2 + 1 / z[-1] ** 2) * uex
    return x, y, z, uex, rhs


@pytest.mark.parametrize("world,mean", [(2, False), (3, False), (4, True)])
def test_poisson_virtual_slabs_reproduce_single_slab(gpu_lib, slab_env, world, mean):
    """BASELINE config 5 (scalar Poisson, copt NDDNDD, u = cos sin sin) through the z-slab path: caller-supplied rhs
    with exchanged halo planes, communication-avoiding smoothing on level 0."""
    from ndsm_b200 import dist as ndist
    x, y, z, uex, rhs = poisson_case((48, 40, 72))
    os.environ["NDSM_VIRTUAL_SLABS"] = "1"
    ref = ndist.poisson_solve(x, y, z, np.zeros_like(uex), rhs, mean=mean)
    os.environ["NDSM_VIRTUAL_SLABS"] = str(world)
    os.environ["NDSM_SLAB_MIN_PLANES"] = "6"
    os.environ["NDSM_SLAB_MIN_POINTS"] = "0"
    got = ndist.poisson_solve(x, y, z, np.zeros_like(uex), rhs, mean=mean)
    assert gpu_lib.ndsm_b200_last_partitioned_levels() > 0
    assert got[0] == ref[0] == 0
    assert np.abs(ref[1] - uex).max() < 2e-2                       # O(h^2)
    if mean:
        assert abs(got[3] - ref[3]) <= 1 and rel_err(got[1], ref[1]) <= 1e-12
    else:
        assert got[3] == ref[3] and got[2] == ref[2]
        assert np.array_equal(got[1], ref[1])


def test_poisson_virtual_slabs_zero_rhs_and_guess(gpu_lib, slab_env):
    from ndsm_b200 import dist as ndist
    x, y, z, uex, _ = poisson_case((40, 36, 60))
    u0 = np.zeros_like(uex)
    u0[:, :, 0] = 0.0
    u0[0], u0[-1], u0[:, 0], u0[:, -1] = 1.0, -0.5, 0.25, 2.0      # Dirichlet data on the D faces (copt NDDNDD)
    os.environ["NDSM_VIRTUAL_SLABS"] = "1"
    ref = ndist.poisson_solve(x, y, z, u0, None)
    os.environ["NDSM_VIRTUAL_SLABS"] = "3"
    os.environ["NDSM_SLAB_MIN_PLANES"] = "6"
    os.environ["NDSM_SLAB_MIN_POINTS"] = "0"
    got = ndist.poisson_solve(x, y, z, u0, None)
    assert gpu_lib.ndsm_b200_last_partitioned_levels() > 0
    assert got[0] == ref[0] == 0 and got[3] == ref[3]
    assert np.array_equal(got[1], ref[1])


def test_poisson_slabs_refuse_unpartitionable_grid(gpu_lib, slab_env, capfd):
    from ndsm_b200 import dist as ndist
    x, y, z, uex, rhs = poisson_case((24, 24, 24))
    os.environ["NDSM_VIRTUAL_SLABS"] = "2"
    os.environ.pop("NDSM_SLAB_MIN_POINTS", None)                   # production threshold: 24^3 stays replicated
    ierr = ndist.poisson_solve(x, y, z, np.zeros_like(uex), rhs)[0]
    assert ierr == 5                                               # NDSM_B200_ERR_ARG
    assert "too small to be partitioned" in capfd.readouterr().err


@pytest.mark.parametrize("world", [2, 3])
def test_poisson_pure_neumann_3d_on_partitioned_levels(gpu_lib, slab_env, oracle, world):
    """All six faces Neumann (copt NNNNNN): the gauge is fixed by subtracting the global mean after every sweep
    (ndsm_optimized.f90:173-189).  On z-partitioned levels the mean is the rank-ordered sum of the slab sums; it
    differs from the single-slab summation order only in rounding, like two OpenMP runs of the reference."""
    from ndsm_b200 import dist as ndist
    nx, ny, nz = 40, 36, 56
    x = np.linspace(0, 1, nx)
    dx = x[1] - x[0]
    y, z = np.arange(ny) * dx, np.arange(nz) * dx
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    uex = np.cos(np.pi * X) * np.cos(2 * np.pi * Y / y[-1]) * np.cos(np.pi * Z / z[-1])   # zero mean, du/dn = 0
    rhs = -(np.pi ** 2) * (1 + 4 / y[-1] ** 2 + 1 / z[-1] ** 2) * uex
    os.environ["NDSM_VIRTUAL_SLABS"] = "1"
    ref = ndist.poisson_solve(x, y, z, np.zeros_like(uex), rhs, copt="NNNNNN")
    os.environ["NDSM_VIRTUAL_SLABS"] = str(world)
    os.environ["NDSM_SLAB_MIN_PLANES"] = "6"
    os.environ["NDSM_SLAB_MIN_POINTS"] = "0"
    got = ndist.poisson_solve(x, y, z, np.zeros_like(uex), rhs, copt="NNNNNN")
    assert gpu_lib.ndsm_b200_last_partitioned_levels() > 0
    assert got[0] == ref[0] == 0
    assert abs(got[3] - ref[3]) <= 1
    assert rel_err(got[1], ref[1]) <= 1e-10
    ora = oracle.poisson_solve([x, y, z], "NNNNNN", np.zeros_like(uex), rhs)
    assert ora[0] == 0 and abs(got[3] - ora[3]) <= 1
    assert rel_err(got[1], ora[1]) <= 1e-10
    assert abs(got[1].mean()) <= 1e-12 * np.abs(got[1]).max() + 1e-15          # the gauge: zero mean
    assert np.abs(got[1] - uex).max() < 5e-2                                    # O(h^2)


@pytest.fixture
def batch_env():
    keys = ("NDSM_BATCH_COMPONENTS", "NDSM_COMPONENT_GROUPS")
    saved = {k: os.environ.get(k) for k in keys}
    yield
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("shape,world,min_planes,groups,kw", [
    ((44, 44, 44), 1, 16, "012", {}),                 # one slab, no communicator
    ((44, 40, 48), 1, 16, "01,2", {}),                # Ax+Ay batched next to Az on a second stream
    ((40, 36, 44), 1, 16, "0,12", {"mean": True}),
    ((40, 33, 52), 3, 4, "012", {}),                  # ragged slabs, two partitioned levels
    ((65, 65, 65), 4, 4, "012", {}),
    ((48, 40, 80), 2, 8, "012", {"mean": True}),      # mean metric: the members' sums are gathered in one message
    ((129, 129, 129), 2, 16, "012", {}),              # the z-lerped prolongation and the direct restriction, batched
])
def test_batched_components_reproduce_member_by_member(gpu_lib, slab_env, batch_env, shape, world, min_planes, groups, kw):
    """Groups of components as ONE launch sequence each (mg_batch.cu) against one solve after the other: the same
    arithmetic per member, so du histories, V-cycle counts, A and B are bit-identical."""
    os.environ.pop("NDSM_BATCH_COMPONENTS", None)
    os.environ.pop("NDSM_COMPONENT_GROUPS", None)
    ref = solve(shape, world, min_planes, **kw)
    assert gpu_lib.ndsm_b200_last_components_mode() == 0
    if groups == "012" and world > 1:
        os.environ["NDSM_BATCH_COMPONENTS"] = "1"     # the short form
    else:
        os.environ["NDSM_COMPONENT_GROUPS"] = groups
    got = solve(shape, world, min_planes, **kw)
    assert gpu_lib.ndsm_b200_last_components_mode() == 2
    again = solve(shape, world, min_planes, **kw)   # cached hierarchies, replayed graphs
    assert gpu_lib.ndsm_b200_last_components_mode() == 2
    for out in (got, again):
        assert out[0] == ref[0] == 0
        for name in ("Ax", "Ay", "Az"):
            assert out[3][name]["du"] == ref[3][name]["du"], name
            assert out[3][name]["nexact"] == ref[3][name]["nexact"], name
        assert np.array_equal(out[1], ref[1])
        assert np.array_equal(out[2], ref[2])
