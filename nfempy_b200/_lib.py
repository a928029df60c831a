"""ctypes binding of libnfa.so (include/nfa.h).  There is no CPU fallback: if the CUDA library is missing the
import of any compute entry point raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFA_LIB", os.path.join(_HERE, "csrc", "libnfa.so"))

_c_i32p = ctypes.POINTER(ctypes.c_int32)
_c_i64p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_d = ctypes.c_double

# name -> (restype, argtypes); every symbol declared in include/nfa.h
SIGNATURES = {
    "nfa_version": (_i, []),
    "nfa_last_error": (ctypes.c_char_p, []),
    "nfa_launch_count": (_i64, []),
    "nfa_timing_enable": (None, [_i]),
    "nfa_timing_read": (_i, [ctypes.POINTER(_d), _c_i64p]),
    "nfa_geom_simplex": (_i, [_i, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "nfa_map_points": (_i, [_i, _i, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nfa_local_h1_elasticity": (
        _i,
        [_i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    ),
    "nfa_local_h1_laplace": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nfa_local_h1_load": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _d, _vp, _vp]),
    "nfa_local_h1_load_pitched": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _d, _vp, _i, _vp]),
    "nfa_local_h1_facet": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _d, _d, _i, _vp, _vp, _vp]),
    "nfa_local_hdiv_l2": (
        _i,
        [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp],
    ),
    "nfa_local_hdiv_facet": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nfa_local_hdiv_l2_general": (
        _i,
        [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64] + [_vp] * 13,
    ),
    "nfa_local_sundus": (
        _i,
        [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64,
         _i64, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    ),
    "nfa_plan_create": (_i, [_i64, _i, ctypes.POINTER(_vp), _c_i64p, _c_i32p, _c_i32p, _vp, ctypes.POINTER(_vp), _vp]),
    "nfa_plan_create_filtered": (_i, [_i64, _i, ctypes.POINTER(_vp), _c_i64p, _c_i32p, _c_i32p, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp]),
    "nfa_filter_local_entries": (_i, [_i64, _i, _vp, _d, _i, _i, _vp, _vp, _vp]),
    "nfa_plan_from_bplan": (_i, [_vp, ctypes.POINTER(_vp), _vp]),
    "nfa_plan_destroy": (None, [_vp]),
    "nfa_plan_nnz": (_i64, [_vp]),
    "nfa_plan_n_dof": (_i64, [_vp]),
    "nfa_plan_bytes": (_i64, [_vp]),
    "nfa_plan_csr": (_i, [_vp, _vp, _vp, _vp]),
    "nfa_plan_reduce": (_i, [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _vp, _i, _vp, _vp]),
    "nfa_csr_spmv": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "nfa_pack_rows": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nfa_unpack_add_rows": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nfa_local_hdiv_coupling": (_i, [_i, _i, _i, _i] + [_vp] * 16 + [_i64] + [_vp] * 5),
    "nfa_local_hcurl_mass": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _d, _vp, _vp, _vp]),
    "nfa_hcurl_field_error": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "nfa_p2p_alloc": (_i, [_i64, ctypes.POINTER(_vp), _vp]),
    "nfa_p2p_open": (_i, [_vp, ctypes.POINTER(_vp)]),
    "nfa_p2p_close": (_i, [_vp]),
    "nfa_p2p_free": (_i, [_vp]),
    "nfa_p2p_timeouts": (_i64, []),
    "nfa_stream_scatter": (_i, [_i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "nfa_rows_copy_i32": (_i, [_i64, _vp, _vp, _vp, _vp, _vp]),
    "nfa_hdiv_rt1_records": (
        _i,
        [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _d,
         _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp],
    ),
    "nfa_rows_locate": (_i, [_i64, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "nfa_rows_expand": (_i, [_i64, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "nfa_plan_from_csr": (_i, [_i64, _i64, _vp, _vp, ctypes.POINTER(_vp)]),
    "nfa_push_rows": (_i, [_vp] * 5 + [_i64] + [_vp] * 5),
    "nfa_pull_add_rows": (_i, [_vp] * 5 + [_i64] + [_vp] * 5),
    "nfa_bplan_gather_load": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "nfa_bplan_n_active_nodes": (_i64, [_vp]),
    "nfa_bplan_load_rhs": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _d, _i, _vp, _vp]),
    "nfa_bplan_create": (_i, [_i64, _i, _i, _vp, _i64, ctypes.POINTER(_vp), _vp]),
    "nfa_bplan_destroy": (None, [_vp]),
    "nfa_bplan_nnz": (_i64, [_vp]),
    "nfa_bplan_bytes": (_i64, [_vp]),
    "nfa_assemble_h1_elasticity_affine": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "nfa_assemble_h1_laplace_affine": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nfa_local_hdiv_tensor": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "nfa_local_hdiv_facet_vec": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _d, _i, _i, _vp, _vp]),
    "nfa_field_error": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "nfa_field_error_scaled": (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i, _i, _vp, _vp,
                                    _vp, _vp, _vp]),
    "nfa_local_hdiv_facet_penalty": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _d, _i, _i, _vp, _vp, _vp]),
    "nfa_local_h1_multifield": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp,
                                     _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "nfa_bplan_block_pattern": (_i, [_vp, _vp, _vp, _vp]),
    "nfa_bplan_n_blocks": (_i64, [_vp]),
    "nfa_bplan_set_p2_items": (_i, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i, _vp]),
    "nfa_p2_lane_order": (_i, [_i64, _vp, _vp, _i, _vp, _vp]),
    "nfa_bplan_has_p2_items": (_i, [_vp]),
    "nfa_bplan_set_p2_split": (_i, [_vp, _i, _vp, _vp]),
    "nfa_bplan_set_cell_mask": (_i, [_vp, _vp, _vp]),
    "nfa_assemble_p2tet_elasticity": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "nfa_assemble_p2tet_elasticity_load": (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i, _vp]),
}

_lib = None


class NfaError(RuntimeError):
    pass


def load():
    """Loads libnfa.so and sets the prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NfaError(
            "libnfa.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C nfempy_b200/csrc` -- there is no CPU fallback" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, where=""):
    if status != 0:
        msg = load().nfa_last_error()
        raise NfaError("%s failed (status %d): %s" % (where or "libnfa call", status, (msg or b"").decode()))


def call(name, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NfaError("libnfa takes device buffers; got a CPU tensor")
    if not t.is_contiguous():
        raise NfaError("libnfa takes contiguous buffers")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """Current torch stream of ``device`` (default: the current device).  libnfa launches on the CURRENT CUDA device:
    callers working on another device wrap their calls in ``torch.cuda.device(device)`` (``GPUAssembler`` does)."""
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(load().nfa_launch_count())
