"""Locate and load the in-tree CUDA library.  There is no CPU fallback: a missing library is an error."""
import ctypes
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "ndsmf.so")
_LIB = None


def load_library(path=None):
    """Return the ctypes handle of ndsmf.so (cached for the default path)."""
    global _LIB
    if path is None:
        if _LIB is not None:
            return _LIB
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "ndsm_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        _LIB = ctypes.CDLL(LIB_PATH)
        _declare(_LIB)
        return _LIB
    lib = ctypes.CDLL(path)
    return lib


def _declare(lib):
    c = ctypes
    vp = c.c_void_p
    lib.ndsm_vector_solve.argtypes = [c.c_size_t] + [vp] * 8
    lib.ndsm_vector_solve.restype = c.c_int
    lib.ndsm_b200_vector_solve_device.argtypes = [vp] * 8
    lib.ndsm_b200_poisson_solve.argtypes = [c.c_int, vp, c.c_char_p, c.c_int, c.c_int, c.c_int, c.c_int,
                                            c.c_double, c.c_double, vp, vp, vp, vp, vp, vp, vp]
    lib.ndsm_b200_poisson_solve_rank.argtypes = [vp, c.c_char_p, c.c_int, c.c_int, c.c_int, c.c_int, c.c_double,
                                                 c.c_double, vp, vp, vp, vp, vp, vp, vp]
    lib.ndsm_b200_last_partitioned_levels.restype = c.c_int
    lib.ndsm_b200_last_components_mode.restype = c.c_int
    lib.ndsm_b200_parse_component_groups.argtypes = [c.c_char_p, vp]
    lib.ndsm_b200_parse_component_groups.restype = c.c_int
    lib.ndsm_b200_new_mg_handle.argtypes = [c.c_int, vp, c.c_int, vp, vp, vp, c.c_int, c.c_int]
    lib.ndsm_b200_new_mg_handle.restype = vp
    lib.ndsm_b200_delete_mg_handle.argtypes = [vp]
    lib.ndsm_b200_delete_mg_handle.restype = None
    lib.ndsm_b200_mg_set_options.argtypes = [vp, c.c_int, c.c_double, c.c_char_p]
    lib.ndsm_b200_mg_ngrids.argtypes = [vp]
    lib.ndsm_b200_mg_level_shape.argtypes = [vp, c.c_int, vp]
    lib.ndsm_b200_mg_level_mesh.argtypes = [vp, c.c_int, c.c_int, vp]
    lib.ndsm_b200_mg_put.argtypes = [vp, c.c_int, c.c_int, vp]
    lib.ndsm_b200_mg_get.argtypes = [vp, c.c_int, c.c_int, vp]
    lib.ndsm_b200_mg_relax.argtypes = [vp, c.c_int, c.c_int]
    lib.ndsm_b200_mg_residual.argtypes = [vp, c.c_int]
    lib.ndsm_b200_mg_restrict.argtypes = [vp, c.c_int]
    lib.ndsm_b200_mg_interp_add.argtypes = [vp, c.c_int]
    lib.ndsm_b200_mg_residual_restrict.argtypes = [vp, c.c_int, vp]
    lib.ndsm_b200_mg_solve_exact.argtypes = [vp, c.c_int, vp]
    lib.ndsm_b200_mg_v_cycle.argtypes = [vp]
    lib.ndsm_b200_mg_solve.argtypes = [vp, c.c_double, c.c_int, vp, vp, vp, vp]
    lib.ndsm_b200_mg_update_u.argtypes = [vp, vp, vp, vp, vp]
    lib.ndsm_b200_plan_create.argtypes = [c.c_int, vp, c.c_int, vp, vp, vp]
    lib.ndsm_b200_plan_create.restype = vp
    lib.ndsm_b200_plan_destroy.argtypes = [vp]
    lib.ndsm_b200_plan_destroy.restype = None
    lib.ndsm_b200_plan_ngrids.argtypes = [vp]
    lib.ndsm_b200_plan_level.argtypes = [vp, c.c_int, vp, vp, vp]
    lib.ndsm_b200_plan_mesh.argtypes = [vp, c.c_int, c.c_int, vp]
    lib.ndsm_b200_plan_interp.argtypes = [vp, c.c_int, c.c_int, vp, vp, vp]
    lib.ndsm_b200_plan_restrict.argtypes = [vp, c.c_int, c.c_int, vp, vp, vp, vp]
    lib.ndsm_b200_ngrids_for.argtypes = [c.c_int]
    lib.ndsm_b200_plan_slab_partition.argtypes = [vp, c.c_int, c.c_int, vp, vp]
    lib.ndsm_b200_plan_sym_heap.argtypes = [c.c_longlong, vp, c.c_int, vp]
    lib.ndsm_b200_bc_setup.argtypes = [vp] * 11
    lib.ndsm_b200_flux_curl.argtypes = [vp, c.c_int, vp, vp, vp, vp, vp, vp]
    lib.ndsm_b200_launch_count.restype = c.c_ulonglong
    lib.ndsm_b200_peer_bytes_sent.restype = c.c_ulonglong
    lib.ndsm_b200_peer_messages_sent.restype = c.c_ulonglong
    lib.ndsm_b200_trace_ncycles.argtypes = [c.c_int]
    lib.ndsm_b200_trace_du.argtypes = [c.c_int, c.c_int]
    lib.ndsm_b200_trace_du.restype = c.c_double
    lib.ndsm_b200_trace_nexact.argtypes = [c.c_int, c.c_int]
    lib.ndsm_b200_last_timing.argtypes = [vp]
    lib.ndsm_b200_version.restype = c.c_char_p
    lib.ndsm_b200_last_slab_points.restype = c.c_ulonglong
    lib.ndsm_b200_dist_unique_id.argtypes = [vp]
    lib.ndsm_b200_dist_init.argtypes = [c.c_int, c.c_int, vp]
    lib.ndsm_b200_dist_transport.restype = c.c_char_p
    lib.ndsm_b200_slab_range.argtypes = [c.c_int, c.c_int, c.c_int, vp, vp]
    lib.ndsm_b200_vector_solve_rank.argtypes = [vp, vp, vp, vp, vp, vp, vp, c.c_int, vp, vp, c.c_int]
    lib.ndsm_b200_release_workspace.restype = None
    lib.ndsm_b200_workspace_bytes.restype = c.c_ulonglong
    lib.ndsm_b200_profile_enable.argtypes = [c.c_int]
    lib.ndsm_b200_profile_enable.restype = None
    lib.ndsm_b200_profile_get.argtypes = [c.c_int, vp, vp]
