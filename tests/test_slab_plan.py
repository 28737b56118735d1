"""Host-side logic of the multi-GPU z-slab decomposition (no GPU needed): partition of every partitioned
level, replication threshold, and the halo property that restriction / prolongation stencils of the
planes a rank produces stay inside its owned planes + NDSM_HALO halo planes."""
import numpy as np
import pytest

from conftest import aniso_mesh

import os

HALO = 6


@pytest.fixture(params=[None, 4, 10])
def halo(request):
    """Halo depth: the compiled default (6 planes) or the NDSM_HALO_PLANES override (read at planning time)."""
    saved = os.environ.get("NDSM_HALO_PLANES")
    if request.param is None:
        os.environ.pop("NDSM_HALO_PLANES", None)
    else:
        os.environ["NDSM_HALO_PLANES"] = str(request.param)
    yield HALO if request.param is None else request.param
    if saved is None:
        os.environ.pop("NDSM_HALO_PLANES", None)
    else:
        os.environ["NDSM_HALO_PLANES"] = saved


@pytest.mark.parametrize("shape,world,min_planes", [((513, 513, 513), 8, 16), ((513, 513, 513), 2, 16),
                                                    ((1025, 1025, 257), 8, 16), ((129, 129, 129), 4, 16),
                                                    ((44, 44, 44), 3, 4), ((40, 33, 52), 2, 4), ((257, 257, 257), 8, 8)])
def test_partition_tiles_and_respects_halo(shape, world, min_planes, halo):
    from ndsm_b200.mg import Plan
    HALO = halo
    min_planes = max(min_planes, HALO)  # a slab must be able to fill its neighbour's halo
    p = Plan(aniso_mesh(shape))
    ndist, zs = p.slab_partition(world, min_planes)
    assert 0 <= ndist <= p.ngrids - 1  # the coarsest level is never partitioned
    if ndist == 0:
        return
    assert zs.shape == (ndist + 1, world + 1)
    for lv in range(ndist + 1):
        nz = p.level(lv)["shape"][2]
        assert zs[lv][0] == 0 and zs[lv][-1] == nz
        assert np.all(np.diff(zs[lv]) >= (min_planes if lv < ndist else 0))
    # finest level is the balanced split used for the output ranges
    assert list(zs[0]) == [shape[2] * r // world for r in range(world + 1)]
    for lv in range(ndist):
        first, count, _, _ = p.restrict_table(lv, 2)
        lo, _, _ = p.interp_table(lv, 2)
        nzf, nzc = p.level(lv)["shape"][2], p.level(lv + 1)["shape"][2]
        for r in range(world):
            f0, f1 = max(zs[lv][r] - (HALO - 1), 0), min(zs[lv][r + 1] + (HALO - 1), nzf)
            for c in range(zs[lv + 1][r], zs[lv + 1][r + 1]):
                assert f0 <= first[c] and first[c] + count[c] <= f1
            if lv + 1 < ndist:
                c0, c1 = zs[lv + 1][r] - HALO, zs[lv + 1][r + 1] + HALO
                for k in range(zs[lv][r], zs[lv][r + 1]):
                    assert c0 <= lo[k] and min(lo[k] + 1, nzc - 1) < c1
    p.close()


def test_expected_depth_for_headline_config():
    """513^3 with the plane-count rule only: on 8 ranks 513 -> 256 -> 128 planes keep >= 16 planes per rank."""
    from ndsm_b200.mg import Plan
    p = Plan(aniso_mesh((513, 513, 513)))
    ndist, zs = p.slab_partition(8, 16)
    assert ndist == 3
    ndist2, _ = p.slab_partition(2, 16)
    assert ndist2 == 5
    ndist1, zs1 = p.slab_partition(1, 16)
    assert ndist1 == 0 and len(zs1) == 0
    p.close()


@pytest.mark.parametrize("shape,world,ndist_expected", [
    ((513, 513, 513), 2, 2),      # 513^3 and 257^3 partitioned, 129^3 (2.1e6 points < 8e6) replicated
    ((513, 513, 513), 4, 2),
    ((513, 513, 513), 8, 3),      # from 8 ranks on the 129^3 level is partitioned too (16 planes per rank)
    ((513, 513, 513), 16, 2),     # 129 / 16 = 8 planes per rank: below the 16-plane minimum
    ((1025, 1025, 257), 8, 2),    # config 4: 257 and 129 planes; 65 / 8 = 8 planes: replicated
    ((129, 129, 129), 8, 1),
    ((129, 129, 129), 4, 0),      # 2.1e6 points: not worth exchanging on four ranks
])
def test_default_partition_depth(shape, world, ndist_expected):
    """The thresholds a solve uses (min_planes < 0 in ndsm_b200_plan_slab_partition): measured choices, DESIGN.md 5."""
    from ndsm_b200.mg import Plan
    saved = {k: os.environ.pop(k, None) for k in ("NDSM_SLAB_MIN_PLANES", "NDSM_SLAB_MIN_POINTS", "NDSM_HALO_PLANES")}
    try:
        p = Plan(aniso_mesh(shape))
        ndist, zs = p.slab_partition(world, -1)
        assert ndist == ndist_expected
        os.environ["NDSM_SLAB_MIN_POINTS"] = "0"       # the override is honoured
        nd0, _ = p.slab_partition(world, -1)
        assert nd0 >= ndist
        p.close()
    finally:
        os.environ.pop("NDSM_SLAB_MIN_POINTS", None)
        for k, v in saved.items():
            if v is not None:
                os.environ[k] = v
