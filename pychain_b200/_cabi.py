"""ctypes binding of libchain_b200.so -- the declarations mirror include/chain_b200.h.

Loading fails loudly: there is no CPU or PyTorch fallback behind this module.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHAIN_B200_LIB") or os.path.join(_PKG, "lib", "libchain_b200.so")

c_int, c_uint, c_float, c_size_t, c_void_p, c_int64 = (
    ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int64)

GRAPH_SHARED, X_IS_EXP, CLAMP_X, GRAD_AS_LOG, LOG_DOMAIN = 1, 2, 4, 8, 16
OK, ERR_INVALID_ARG, ERR_WORKSPACE, ERR_UNSUPPORTED, ERR_BAD_BLOB = 0, -1, -2, -3, -4

# every symbol include/chain_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "chain_b200_version": (c_int, []),
    "chain_b200_status_string": (ctypes.c_char_p, [c_int]),
    "chain_b200_set_verbose_level": (None, [c_int]),
    "chain_b200_get_verbose_level": (c_int, []),
    "chain_b200_last_launch_count": (c_int, []),
    "chain_b200_workspace_bytes": (c_size_t, [c_int] * 5 + [c_uint]),
    "chain_b200_forward_backward": (
        c_int, [c_void_p] * 11 + [c_int] * 5 + [c_float, c_uint, c_float] + [c_void_p] * 5 + [c_void_p, c_size_t, c_int, c_void_p]),
    "chain_b200_forward_backward_log_domain": (
        c_int, [c_void_p] * 10 + [c_int] * 5 + [c_uint, c_float] + [c_void_p] * 5 + [c_void_p, c_size_t, c_int, c_void_p]),
    "chain_b200_den_prepare_bytes": (c_int64, [c_void_p] * 4 + [c_int] * 3 + [c_uint]),
    "chain_b200_den_prepare": (c_int, [c_void_p] * 9 + [c_int] * 3 + [c_uint, c_void_p, c_int64]),
    "chain_b200_den_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "chain_b200_den_group_width": (c_int, [c_int] * 4),
    "chain_b200_den_forward_backward": (
        c_int, [c_void_p] * 2 + [c_int] + [c_void_p] * 2 + [c_int] * 3 + [c_float, c_uint, c_float] + [c_void_p] * 5 +
        [c_void_p, c_size_t, c_int, c_void_p]),
    "chain_b200_combine_loss": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "chain_b200_scale_grad": (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
}

_lib = None


class ChainB200Error(RuntimeError):
    pass


def lib():
    """The loaded library (loads on first use).  Raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ChainB200Error(
                "libchain_b200.so is not built ({}); run `python -m pychain_b200.build`. "
                "pychain_b200 has no CPU / PyTorch fallback.".format(LIB_PATH))
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().chain_b200_status_string(status).decode()
        raise ChainB200Error("{} failed: {} (status {})".format(what, msg, status))
