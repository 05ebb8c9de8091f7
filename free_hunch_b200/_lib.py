"""ctypes binding of libfreehunch_b200.so (C ABI: include/free_hunch_b200.h).

There is no CPU fallback: the only library this loader opens is the sm_100a build next to this file, and every
entry point fails with FH_ECUDA without a CUDA device.  If the library is missing, first use fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfreehunch_b200.so")

c_void_p, c_int, c_double, c_long = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_longlong

_SIGNATURES = {
    "fh_last_error": (ctypes.c_char_p, []),
    "fh_version": (c_int, []),
    "fh_launch_count": (c_long, [c_int]),
    "fh_cg_kernel_time": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.POINTER(ctypes.c_float), c_void_p]),
    "fh_cg_solve_async": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p]),
    "fh_cg_solve_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "fh_cg_solve_f64_async": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p]),
    "fh_cg_resident_time": (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.POINTER(ctypes.c_float), c_void_p]),
    "fh_trace_read": (c_int, [c_void_p, c_void_p]),
    "fh_prof_enable": (c_int, [c_int]),
    "fh_prof_report": (c_int, [ctypes.c_char_p, c_int]),
    "fh_ctx_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int]),
    "fh_ctx_destroy": (c_int, [c_void_p]),
    "fh_ctx_set_cg_mode": (c_int, [c_void_p, c_int]),
    "fh_ctx_cg_last_mode": (c_int, [c_void_p]),
    "fh_ctx_fused_updates": (c_int, [c_void_p]),
    "fh_dct2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "fh_state_create": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_int, c_int]),
    "fh_state_destroy": (c_int, [c_void_p]),
    "fh_state_init": (c_int, [c_void_p, c_void_p, c_double, c_void_p]),
    "fh_state_rank": (c_int, [c_void_p]),
    "fh_state_var2": (c_double, [c_void_p]),
    "fh_cov_dot": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "fh_cov_dot_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "fh_time_update": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "fh_time_update_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "fh_time_update_from_mean": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "fh_mech_combine": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double, c_void_p, c_void_p, c_int, c_void_p]),
    "fh_mech_flag": (c_int, [c_void_p, c_void_p, c_double, c_double, c_void_p, c_int, c_void_p]),
    "fh_cov_dot_f64_if": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fh_mech_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_void_p, c_void_p, c_int, c_void_p]),
    "fh_mech_allclose": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "fh_sampler_euler": (c_int, [c_void_p, c_void_p, c_double, c_double, c_void_p, c_void_p, c_long, c_void_p]),
    "fh_sampler_heun": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_void_p, c_long, c_void_p]),
    "fh_space_update_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_int, c_void_p]),
    "fh_space_update": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_int, c_void_p]),
    "fh_state_set_cov_vectors": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_double, c_void_p]),
    "fh_state_drop_all": (c_int, [c_void_p, c_double, c_void_p]),
    "fh_state_gram": (c_int, [c_void_p, c_void_p, c_void_p]),
    "fh_state_recombine": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "fh_state_export": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fh_op_create_conv": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_void_p, c_int, c_int]),
    "fh_op_create_inpaint": (c_int, [c_void_p, ctypes.POINTER(c_void_p), c_void_p, c_int]),
    "fh_op_destroy": (c_int, [c_void_p]),
    "fh_op_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "fh_cg_solve": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_int, c_int, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class FreeHunchLibraryError(RuntimeError):
    pass


def _bind(path):
    if not os.path.exists(path):
        raise FreeHunchLibraryError(
            f"{path} not found: build the sm_100a extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "free_hunch_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    """The loaded library (loads on first use)."""
    global _lib
    if _lib is None:
        _lib = _bind(_LIB_PATH)
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().fh_last_error()
        raise FreeHunchLibraryError(f"libfreehunch_b200 error {rc}: {msg.decode() if msg else '?'}")
