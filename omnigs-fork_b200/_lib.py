"""ctypes binding of libomnigs_b200.so (the C ABI declared in include/omnigs_b200.h).

There is no fallback: if the CUDA library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libomnigs_b200.so"
_lib = None


class OgsError(RuntimeError):
    """A libomnigs_b200 call returned a non-zero status."""


def library_path():
    # OMNIGS_B200_LIB: an alternative build of the same library (A/B measurements of kernel variants)
    return os.environ.get("OMNIGS_B200_LIB") or os.path.join(_HERE, _LIB_NAME)


# every symbol include/omnigs_b200.h declares: name -> (restype, argtypes)
_c_int, _c_i64, _c_sz, _c_f = ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_float
_p = ctypes.c_void_p
SYMBOLS = {
    "ogs_abi_version": (_c_int, []),
    "ogs_last_error": (ctypes.c_char_p, []),
    "ogs_geom_bytes": (_c_sz, [_c_int]),
    "ogs_img_bytes": (_c_sz, [_c_int, _c_int]),
    "ogs_binning_bytes": (_c_sz, [_c_i64, _c_int, _c_int]),
    "ogs_lonlat_forward_stage1": (_c_int, [_c_int] * 5 + [_p] * 5 + [_c_f] + [_p] * 4 + [_p] * 3 + [ctypes.POINTER(_c_i64), _p]),
    "ogs_lonlat_forward_stage1_band": (_c_int, [_c_int] * 7 + [_p] * 5 + [_c_f] + [_p] * 4 + [_p] * 3 + [ctypes.POINTER(_c_i64), _p]),
    "ogs_lonlat_forward_stage2": (_c_int, [_c_int] * 3 + [_c_i64] + [_p] * 6),
    "ogs_lonlat_backward": (_c_int, [_c_int] * 3 + [_c_i64, _c_int, _c_int] + [_p] * 5 + [_c_f] + [_p] * 5 + [_p] * 3 + [_p] + [_p] * 9 + [_p]),
    "ogs_lonlat_backward_render": (_c_int, [_c_int, _c_i64, _c_int, _c_int] + [_p] * 5 + [_p]),
    "ogs_lonlat_backward_finish": (_c_int, [_c_int] * 5 + [_p] * 3 + [_c_f] + [_p] * 6 + [_p] * 9 + [_p]),
    "ogs_grad_acc_offset": (_c_sz, [_c_int]),
    "ogs_mark_all_visible": (_c_int, [_c_int, _p, _p]),
    "ogs_export_geometry": (_c_int, [_c_int, _p] + [_p] * 7 + [_p]),
    "ogs_export_binning": (_c_int, [_c_int] * 3 + [_c_i64] + [_p] * 3 + [_p] * 5 + [_p]),
    "ogs_pinhole_forward_stage1": (_c_int, [_c_int] * 5 + [_p] * 5 + [_c_f] + [_p] * 5 + [_c_f] * 2 + [_c_int] + [_p] * 3
                                   + [ctypes.POINTER(_c_i64), _p]),
    "ogs_pinhole_backward": (_c_int, [_c_int] * 3 + [_c_i64, _c_int, _c_int] + [_p] * 5 + [_c_f] + [_p] * 5 + [_c_f] * 2
                             + [_p] * 4 + [_p] + [_p] * 9 + [_p]),
    "ogs_mark_visible_pinhole": (_c_int, [_c_int] + [_p] * 4 + [_p]),
    "ogs_lonlat_forward_raw_stage1": (_c_int, [_c_int] * 5 + [_p] * 5 + [_c_f] + [_p] * 3 + [_p] * 3
                                      + [ctypes.POINTER(_c_i64), _p]),
    "ogs_lonlat_backward_raw": (_c_int, [_c_int] * 3 + [_c_i64] + [_c_int] * 2 + [_p] + [_p] * 4 + [_c_f] + [_p] * 4
                                + [_p] * 3 + [_p] + [_p] * 7 + [_p]),
    "ogs_photometric_loss_workspace_bytes": (_c_sz, [_c_int, _c_int]),
    "ogs_photometric_loss": (_c_int, [_c_int] * 3 + [_c_f] + [_p] * 3 + [_c_int] + [_p] * 3 + [_p]),
    "ogs_adam_step": (_c_int, [_c_int] + [_p] * 6 + [_c_i64] + [ctypes.c_double] * 3 + [_p]),
    "ogs_densify_stats": (_c_int, [_c_int] + [_p] * 5 + [_p]),
    "ogs_view_stats": (_c_int, [_c_int] + [_p] * 5 + [_p]),
    "ogs_set_seam_wrap": (_c_int, [_c_int]),
    "ogs_get_seam_wrap": (_c_int, []),
    "ogs_profile_enable": (_c_int, [_c_int]),
    "ogs_profile_read": (_c_int, [ctypes.POINTER(_c_f), _c_int]),
    "ogs_lonlat_forward_bin": (_c_int, [_c_int] * 3 + [_c_i64] + [_p] * 3 + [_p]),
    "ogs_lonlat_forward_blend": (_c_int, [_c_int] * 3 + [_c_i64] + [_p] * 5 + [_p]),
    "ogs_lonlat_forward_stage1_geometry": (_c_int, [_c_int] * 3 + [_p] * 3 + [_c_f] + [_p] * 4 + [_p] * 3 + [ctypes.POINTER(_c_i64), _p]),
    "ogs_lonlat_forward_colors": (_c_int, [_c_int] * 3 + [_p] * 5 + [_p]),
    "ogs_lonlat_backward_render_into": (_c_int, [_c_int, _c_i64, _c_int, _c_int] + [_p] * 6 + [_p]),
    "ogs_lonlat_backward_finish_from": (_c_int, [_c_int] * 5 + [_p] * 3 + [_c_f] + [_p] * 6 + [_p] + [_p] * 9 + [_p]),
    "ogs_lonlat_backward_finish_range": (_c_int, [_c_int] * 5 + [_p] * 3 + [_c_f] + [_p] * 6 + [_p, _c_int, _c_int] + [_p] * 9 + [_p]),
    "ogs_lonlat_backward_view": (_c_int, [_c_int] * 3 + [_c_i64, _c_int, _c_int] + [_p] * 4 + [_c_f] + [_p] * 4 + [_p] * 4
                                 + [_c_int] + [_p] * 9 + [_p]),
    "ogs_sh_gradient_from_views": (_c_int, [_c_int] * 4 + [_p] * 6 + [_p]),
    "ogs_export_pair_counts": (_c_int, [_c_int] * 3 + [_c_i64] + [_p] * 4 + [_p]),
    "ogs_band_rows_allgather": (_c_int, [_p, _c_int, _c_int, _p] + [_c_int] * 4 + [_p]),
    "ogs_peer_allreduce": (_c_int, [_p, _c_int, _c_int, _c_sz, _c_sz, _p]),
    "ogs_multimem_allreduce": (_c_int, [_p, _c_int, _c_int, _c_sz, _c_sz, _p]),
}


def load_library():
    """Load (once) and return the ctypes handle; raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise OgsError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the rasterizer)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    if lib.ogs_abi_version() != 2:
        raise OgsError(f"ABI version mismatch: library reports {lib.ogs_abi_version()}, host expects 2")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load_library().ogs_last_error().decode("utf-8", "replace")
        raise OgsError(f"libomnigs_b200 error {status}: {msg}")
