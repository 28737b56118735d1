"""CPU-side checks of the product library: every symbol the header declares is exported, the reference
getters return the reference's values, the host-only planner reproduces the oracle's hierarchy and tables,
and -- without a GPU -- compute entry points fail loudly instead of falling back."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

from conftest import REPO, aniso_mesh

HEADER = os.path.join(REPO, "include", "ndsm_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n.startswith(("ndsm_", "get_"))))


@pytest.fixture(scope="module")
def lib():
    from ndsm_b200 import load_library
    return load_library()


def test_header_symbols_are_exported(lib):
    names = declared_functions()
    assert "ndsm_vector_solve" in names and "get_iopt_iopt_nmaxex" in names and len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_reference_getters(lib):
    """Values of ndsm_vector_potential.f90:40-57 as returned by ndsm_python_wrapper.f90:164-234."""
    want = {"get_iopt_len": 16, "get_iopt_ierr": 16, "get_iopt_ms": 0, "get_iopt_ncycles": 1, "get_iopt_debug": 5,
            "get_iopt_dumax": 6, "get_iopt_iopt_nmaxex": 7, "get_iopt_true": 1, "get_iopt_false": 0,
            "get_ropt_tim": 2, "get_ropt_vtol": 0, "get_ropt_ctol": 1}
    for k, v in want.items():
        assert getattr(lib, k)() == v, k


def test_mirror_signature_matches_reference():
    """Host mirror keeps the reference's argument names, order and defaults (ndsm.py:66)."""
    from ndsm_b200 import vector_potential
    sig = inspect.signature(vector_potential)
    names = list(sig.parameters)
    ref = ["x", "y", "z", "b", "niterex_max", "ncycles_max", "ex_tol", "vc_tol", "ms", "mean", "libname", "libpath",
           "debug"]
    assert names[: len(ref)] == ref
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["niterex_max"], d["ncycles_max"], d["ex_tol"], d["vc_tol"], d["ms"], d["mean"], d["libname"],
            d["libpath"], d["debug"]) == (10000, 1024, 1e-13, 1e-10, 5, False, "ndsmf.so", None, False)
    ref_py = "/root/reference/ndsm.py"
    if os.path.exists(ref_py):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_ndsm", ref_py)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        rsig = inspect.signature(mod.vector_potential)
        assert list(rsig.parameters) == ref
        assert {k: v.default for k, v in rsig.parameters.items()} == {k: d[k] for k in ref}


def test_option_vectors_filled_like_reference(lib):
    from ndsm_b200.ndsm import _options
    ioptc, ropt = _options(lib, 777, 55, 1e-9, 1e-7, 3, True, True, flxcrl=1)
    assert ioptc.dtype == np.intc and ioptc.size == 16 and ropt.size == 16
    assert (ioptc[0], ioptc[1], ioptc[4], ioptc[5], ioptc[6], ioptc[7]) == (3, 55, 1, 1, 0, 777)
    assert (ropt[0], ropt[1]) == (1e-7, 1e-9)


def test_no_gpu_means_loud_failure(lib, capfd):
    if lib.ndsm_b200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from ndsm_b200 import synthetic, vector_potential
    x, y, z = synthetic.mesh(8)
    ierr, A, B = vector_potential(x, y, z, np.ones((3, 8, 8, 8)))
    assert ierr == 3  # NDSM_B200_ERR_CUDA -- never a CPU result
    assert not A.any()
    assert "no CUDA device" in capfd.readouterr().err
    from ndsm_b200.mg import MGHandle
    with pytest.raises(RuntimeError):
        MGHandle(aniso_mesh((8, 8, 8)), "NDDNDD")


@pytest.mark.parametrize("nmin,ng", [(4, 1), (7, 1), (8, 2), (16, 3), (22, 3), (129, 6), (220, 6), (255, 6), (256, 7),
                                     (257, 7), (513, 8), (1025, 9)])
def test_ngrids_rule(lib, oracle, nmin, ng):
    assert lib.ndsm_b200_ngrids_for(nmin) == ng == oracle.lib().orc_ngrids(nmin)


@pytest.mark.parametrize("shape", [(22, 22, 22), (129, 129, 33), (513, 513, 513), (1025, 1025, 257), (65, 18), (513, 257)])
def test_plan_matches_oracle_hierarchy_and_tables(lib, oracle, shape):
    from ndsm_b200.mg import Plan
    mesh = aniso_mesh(shape)
    p = Plan(mesh)
    L = oracle.lib()
    ng = L.orc_ngrids(min(shape))
    assert p.ngrids == ng
    sh = list(shape)
    meshes = [mesh]
    for g in range(ng):
        lv = p.level(g)
        assert list(lv["shape"][: len(shape)]) == sh
        # layout invariants: 64-byte rows, 256-byte planes, colour arrays do not overlap
        assert lv["mcnt"] == (sh[0] + 1) // 2 and lv["hp"] % 8 == 0 and lv["hp"] >= lv["mcnt"]
        assert lv["ps"] % 32 == 0 and lv["ps"] >= lv["hp"] * sh[1]
        assert lv["cs"] == lv["ps"] * (sh[2] if len(shape) == 3 else 1)
        if g > 0:  # regenerated coarse meshes (ndsm_multigrid_core.f90:253-259)
            ms = [((np.arange(n) * (m0.max() - m0.min())) / (n - 1) + m0.min()) for n, m0 in zip(sh, mesh)]
            for d in range(len(shape)):
                assert np.array_equal(p.mesh_of(g, d), ms[d])
            meshes.append(ms)
        sh = [max(n // 2, 1) for n in sh]
    vp = ctypes.c_void_p
    for g in range(ng - 1):
        for d in range(len(shape)):
            qf, qc = np.ascontiguousarray(meshes[g][d]), np.ascontiguousarray(meshes[g + 1][d])
            lo, wl, wh = p.interp_table(g, d)
            olo = np.zeros(qf.size, dtype=np.int64)
            owl, owh = np.zeros(qf.size), np.zeros(qf.size)
            L.orc_interp_table(qf.size, qf.ctypes.data_as(vp), qc.size, qc.ctypes.data_as(vp), olo.ctypes.data_as(vp),
                               owl.ctypes.data_as(vp), owh.ctypes.data_as(vp))
            assert np.array_equal(lo, olo) and np.array_equal(wl, owl) and np.array_equal(wh, owh)
            assert np.allclose(wl + wh, 1.0, atol=1e-12)
            first, count, c2, w2 = p.restrict_table(g, d)
            ofirst, ocount = np.zeros(qc.size, dtype=np.int64), np.zeros(qc.size, dtype=np.int64)
            oc2 = np.zeros((qc.size, 8))
            ow2 = ctypes.c_double(0)
            wmax = L.orc_restrict_table(qf.size, qf.ctypes.data_as(vp), qc.size, qc.ctypes.data_as(vp),
                                        ofirst.ctypes.data_as(vp), ocount.ctypes.data_as(vp), oc2.ctypes.data_as(vp),
                                        ctypes.byref(ow2))
            assert 3 <= wmax <= 8
            assert np.array_equal(first, ofirst) and np.array_equal(count, ocount)
            assert np.array_equal(c2, oc2) and w2 == ow2.value
    p.close()


def test_plan_rejects_shapes_without_a_hierarchy(lib):
    from ndsm_b200.mg import Plan
    with pytest.raises(ValueError):
        Plan(aniso_mesh((3, 9, 9)))


@pytest.mark.parametrize("spec,ngroups,group_of", [
    ("012", 1, [0, 0, 0]),
    ("210", 1, [0, 0, 0]),
    ("01,2", 2, [0, 0, 1]),
    ("2,01", 2, [0, 0, 1]),          # groups are ordered by their lowest component
    ("02,1", 2, [0, 1, 0]),
    ("0,12", 2, [0, 1, 1]),
    ("0,1,2", 3, [0, 1, 2]),
    ("", 0, [-1, -1, -1]),           # not a partition of {0,1,2}: ignored
    ("01", 0, [-1, -1, -1]),
    ("011,2", 0, [-1, -1, -1]),
    ("0,,12", 0, [-1, -1, -1]),
    ("013", 0, [-1, -1, -1]),
    ("01,2,", 0, [-1, -1, -1]),
    ("0 1 2", 0, [-1, -1, -1]),
])
def test_component_group_specs(lib, spec, ngroups, group_of):
    """NDSM_COMPONENT_GROUPS (which component solves are batched into one launch sequence, mg_batch.cu): only a
    partition of the three components is accepted, anything else leaves the default scheduling in place."""
    out = (ctypes.c_int * 3)(7, 7, 7)
    assert lib.ndsm_b200_parse_component_groups(spec.encode(), out) == ngroups
    assert list(out) == group_of
