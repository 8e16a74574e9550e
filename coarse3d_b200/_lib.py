"""ctypes binding of libcoarse3d_b200.so (the C ABI in include/coarse3d_b200.h).

The library is the product: if it is missing or fails to load, importing any
operator raises -- there is no CPU or PyTorch fallback.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcoarse3d_b200.so")

C3D_OK = 0


class C3DError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "coarse3d_b200: %s is missing. Build it with `python coarse3d_b200/build.py` "
            "(needs nvcc; sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    P = c_void_p
    sigs = {
        "c3d_version": (c_int, []),
        "c3d_last_error": (c_char_p, []),
        "c3d_launch_count": (c_longlong, []),
        "c3d_project_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
        "c3d_project_cluster_supported": (c_int, [c_int, c_int, c_int]),
        "c3d_project_batch": (c_int, [P, c_int, P, c_int, c_int64, P, c_double, c_double, c_double,
                                      c_double, c_int, c_int, P, P, P, P, P, P, P, P, c_int, P, P, c_size_t, P]),
        "c3d_project_assemble_batch": (c_int, [P, P, c_int, c_int64, P, P, P, c_int, P, P, c_double, c_double,
                                               c_double, c_double, c_int, c_int, P, P, P, P, P, P, P, P,
                                               P, c_int, P, P, c_size_t, P]),
        "c3d_unproject_confusion_batch": (c_int, [P, P, P, P, P, c_int, c_int64, c_int, c_int, c_int, c_int,
                                                  c_int, c_int, P, P, P, P]),
        "c3d_entropy_select_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
        "c3d_entropy_select_batch": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P,
                                             c_uint64, P, P, P, P]),
        "c3d_lovasz_workspace_bytes": (c_size_t, [c_int, c_int64]),
        "c3d_lovasz_forward": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_uint64, c_int64, P, P, P]),
        "c3d_lovasz_backward": (c_int, [c_int, c_int, c_int, c_int, c_int, c_uint64, c_int64, P, P, P, c_int, P]),
        "c3d_lovasz_info": (c_int, [P, P, P]),
        "c3d_knn_batch": (c_int, [P, P, P, P, P, P, c_int, c_int64, c_int, c_int, c_int, c_int,
                                  c_float, c_int, P, c_int, c_int, P, P, c_size_t, P]),
        "c3d_profile_enable": (c_int, [c_char_p]),
        "c3d_profile_read": (c_int, [c_char_p, P, P]),
        "c3d_profile_names": (c_int, [P, c_int]),
        "c3d_profile_timeline": (c_int, [P, c_int]),
        "c3d_profile_reset": (c_int, []),
        "c3d_proto_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
        "c3d_proto_loss_forward": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int,
                                           c_int, c_float, c_float, c_int, P, c_int, c_uint64, c_int, P,
                                           P, P]),
        "c3d_proto_loss_forward_phase": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int,
                                                 c_int, c_float, c_float, c_int, P, c_int, c_uint64, c_int,
                                                 c_int, P, P, P]),
        "c3d_proto_loss_backward": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P,
                                            c_int, P]),
        "c3d_zero_fill": (c_int, [P, c_size_t, P]),
        "c3d_zero_fill_daemon": (c_int, [P, c_size_t, c_int, c_int, c_int, c_int, P, P, P]),
        "c3d_delay": (c_int, [c_uint64, P]),
        "c3d_zero_fill_background": (c_int, [P, c_size_t, c_int, c_int, c_int, c_int, P]),
        "c3d_proto_loss_info": (c_int, [P, P, P]),
        "c3d_proto_loss_rows": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, P, P, P, P]),
        "c3d_proto_ema_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int64]),
        "c3d_proto_ema_accumulate": (c_int, [P, P, P, P, P, P, P, c_float, c_int, c_int, c_int, c_int,
                                             c_int, c_int, c_int, c_int64, P, c_int, c_uint64, P, P, P, P]),
        "c3d_proto_ema_accumulate_dense": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                   c_int64, P, c_int, c_uint64, P, P, P, P]),
        "c3d_proto_ema_apply": (c_int, [P, P, c_int, c_int, c_int, c_int, c_double, P, P, P, P]),
        "c3d_proto_ema_info": (c_int, [P, P, P]),
        "c3d_proto_bank_normalise": (c_int, [P, c_int, c_int, P, P]),
        "c3d_knn_sort_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int]),
        "c3d_knn_sort_points": (c_int, [P, P, P, P, c_int, c_int64, c_int, c_int, c_int, P, P, P]),
        "c3d_peer_exchange_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
        "c3d_peer_state_bytes": (c_size_t, [c_int, c_int]),
        "c3d_peer_alloc": (c_int, [c_size_t, P]),
        "c3d_peer_free": (c_int, [P]),
        "c3d_peer_export": (c_int, [P, P]),
        "c3d_peer_import": (c_int, [P, P]),
        "c3d_peer_close": (c_int, [P]),
        "c3d_proto_ema_apply_peers": (c_int, [P, P, P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_double,
                                              P, P, P, c_double, P]),
        "c3d_proto_step_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int64]),
        "c3d_proto_step": (c_int, [P, P, P, P, P, P, P, P, P, c_float, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_float, c_float, c_int, P, c_int, P, c_int, c_uint64, c_int64, c_int,
                                   c_int, P, P, P, P, P, P, P, c_size_t, P]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib, sigs


lib, SIGNATURES = _load()


def check(status):
    if status != C3D_OK:
        msg = lib.c3d_last_error().decode("utf-8", "replace")
        if status == 1:
            raise ValueError(msg)
        raise C3DError("coarse3d_b200 status %d: %s" % (status, msg))


def launch_count():
    return int(lib.c3d_launch_count())
