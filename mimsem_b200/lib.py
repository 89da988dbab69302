"""ctypes loader for libmimsem_gpu.so -- fails loudly when the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MIMSEM_GPU_LIB") or os.path.join(_HERE, "libmimsem_gpu.so")   # (override: diagnostics builds, scripts/)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)
_vp = C.c_void_p

# every symbol include/mimsem_gpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "mimsem_last_error": (C.c_char_p, []),
    "mimsem_basis_gll": (C.c_int, [C.c_int, _dp, _dp]),
    "mimsem_basis_tables": (C.c_int, [C.c_int, C.c_int, _dp, _dp]),
    "mimsem_elmat": (C.c_int, [C.c_int, C.c_int, C.c_int, _dp]),
    "mimsem_topo_patch_sizes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip]),
    "mimsem_topo_patch": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _ip]),
    "mimsem_topo_write_input": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]),
    "mimsem_mesh_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "mimsem_mesh_destroy": (None, [_vp]),
    "mimsem_mesh_sizes": (C.c_int, [_vp, _lp]),
    "mimsem_mesh_tables": (C.c_int, [_vp, _ip, _ip, _ip, _ip, _ip]),
    "mimsem_mesh_geometry": (C.c_int, [_vp, _dp, _dp]),
    "mimsem_mesh_coords": (C.c_int, [_vp, _dp]),
    "mimsem_gpu_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "mimsem_gpu_destroy": (C.c_int, [_vp]),
    "mimsem_gpu_set_option": (C.c_int, [_vp, C.c_char_p, C.c_longlong]),
    "mimsem_gpu_set_basis": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp]),
    "mimsem_gpu_set_topo": (C.c_int, [_vp] + [C.c_int] * 7 + [_ip] * 5),
    "mimsem_gpu_set_geom": (C.c_int, [_vp, _dp, _dp]),
    "mimsem_gpu_set_element_keys": (C.c_int, [_vp, C.c_int, _ip]),
    "mimsem_gpu_set_ghosts": (C.c_int, [_vp, C.c_int, C.c_int, _ip]),
    "mimsem_gpu_set_thickness": (C.c_int, [_vp, C.c_int, _dp]),
    "mimsem_gpu_sizes": (C.c_int, [_vp, _lp]),
    "mimsem_gpu_levels_to_columns": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_columns_to_levels": (C.c_int, [_vp, C.c_int, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_form_permutation": (C.c_int, [_vp, C.c_int, _ip]),
    "mimsem_gpu_apply_M1": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_apply_M2": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_apply_M0": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_apply_M1h": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_apply_M2h": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_apply_M0h": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_apply_K": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_apply_R": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_apply_R_up": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, C.c_double, _vp, _vp, _vp]),
    "mimsem_gpu_apply_M0h_up": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, C.c_double, _vp, _vp, _vp]),
    "mimsem_gpu_solve_M1": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, C.c_double, C.c_int,
                                      _ip, _dp, _vp]),
    "mimsem_gpu_solve_M0": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_diag_M1": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp]),
    "mimsem_gpu_diag_M0": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_solve_M2": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_apply_UtQW": (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_columns_to_vertical": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_vertical_to_columns": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_pc_bjacobi_M1": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_apply_incidence": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_incidence_csr": (C.c_int, [_vp, C.c_int, _lp, _lp, _ip, _dp]),
    "mimsem_gpu_apply_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp]),
    "mimsem_gpu_apply_host_up": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _dp, _dp, C.c_double, _dp, _dp]),
    "mimsem_gpu_gather_rows": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_scatter_rows": (C.c_int, [_vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_ipc_alloc": (C.c_int, [_vp, C.c_int64, C.POINTER(_vp), C.c_char_p]),
    "mimsem_gpu_ipc_open": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "mimsem_gpu_ipc_close": (C.c_int, [_vp, _vp, C.c_int]),
    "mimsem_gpu_halo_push": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_halo_pull": (C.c_int, [_vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_apply_M1_halo": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int,
                                           C.c_int, _vp, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_apply_M1_halo_ll": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int,
                                              C.c_int, _vp, C.c_int, _vp, _vp, C.c_int64, C.c_int, C.c_int, _vp, _vp, _vp]),
    "mimsem_gpu_solve_M1_dist": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _vp, _vp, C.c_double, C.c_int,
                                           _ip, _dp, _vp, _vp, _vp]),
    "mimsem_gpu_launch_count": (C.c_int64, [_vp]),
    "mimsem_gpu_apply_M1ray": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "mimsem_gpu_dev_alloc": (C.c_int, [_vp, C.c_int64, C.POINTER(_vp)]),
    "mimsem_gpu_dev_free": (C.c_int, [_vp, _vp]),
    "mimsem_gpu_stream_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "mimsem_gpu_stream_destroy": (C.c_int, [_vp, _vp]),
    "mimsem_gpu_graph_begin": (C.c_int, [_vp, _vp]),
    "mimsem_gpu_graph_end": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "mimsem_gpu_graph_launch": (C.c_int, [_vp, _vp, _vp]),
    "mimsem_gpu_graph_destroy": (C.c_int, [_vp, _vp]),
    "mimsem_gpu_host_alloc": (C.c_int, [_vp, C.c_int64, C.POINTER(_vp)]),
    "mimsem_gpu_host_free": (C.c_int, [_vp, _vp]),
    "mimsem_gpu_dev_copy": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int]),
    "mimsem_gpu_dev_sync": (C.c_int, [_vp, _vp]),
}


class HaloDesc(C.Structure):
    """mimsem_halo_desc (include/mimsem_gpu.h)"""
    _fields_ = [("npush", C.c_int), ("d_push", C.c_void_p), ("npull", C.c_int), ("d_pull", C.c_void_p), ("d_inbox", C.c_void_p),
                ("stride", C.c_int64), ("nbuf", C.c_int), ("push_ctas", C.c_int), ("d_epoch", C.c_void_p), ("d_err", C.c_void_p),
                ("ll", C.c_int)]


class ReduceDesc(C.Structure):
    """mimsem_reduce_desc (include/mimsem_gpu.h)"""
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("d_peer_areas", C.c_void_p), ("d_seq", C.c_void_p), ("d_err", C.c_void_p)]


class MimsemError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load libmimsem_gpu.so (built by mimsem_b200/csrc/Makefile or __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MimsemError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise MimsemError("mimsem error %d: %s" % (rc, load_library().mimsem_last_error().decode()))
