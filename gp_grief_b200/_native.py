"""ctypes binding of libgrief_b200.so (C ABI: include/grief_b200.h).

There is NO fallback: if the shared library is missing or fails to load, every entry point raises.
PyTorch is used only to own device buffers and streams; the signatures below carry raw pointers.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libgrief_b200.so")

OK, ERR_BAD_ARG, ERR_CUDA, ERR_NOT_PD, ERR_UNSUPPORTED, ERR_LIBRARY = 0, 1, 2, 3, 4, 5
SC_LML, SC_YT_ALPHA, SC_LOGDET, SC_GRAD_NOISE, SC_RTB, SC_ALPHA_SQ, SC_TRACE, SC_COUNT = range(8)
KERNEL_IDS = {"RBF": 0, "Exponential": 1, "Matern32": 2, "Matern52": 3, "host": 4}
OPT_GEMM_MODE, OPT_DIGITS_GRAM, OPT_DIGITS_Z, OPT_SLAB_BUDGET, OPT_DIGITS_VAR = 0, 1, 2, 3, 4

c_int, c_i64, c_size, c_void, c_dbl = ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_double
_P = ctypes.POINTER

# name -> (restype, argtypes): every symbol include/grief_b200.h declares
SIGNATURES = {
    "grief_version": (c_int, []),
    "grief_last_error": (ctypes.c_char_p, []),
    "grief_launch_count": (c_int, []),
    "grief_launch_count_reset": (None, []),
    "grief_profile_enable": (None, [c_int]),
    "grief_profile_slots": (c_int, []),
    "grief_profile_read": (None, [c_void, c_void]),
    "grief_ctx_create": (c_int, [_P(c_void)]),
    "grief_ctx_destroy": (None, [c_void]),
    "grief_topk_kron": (c_int, [c_int, c_void, c_void, c_void, c_int, c_void, c_void, _P(c_int), c_void]),
    "grief_plan_create": (c_int, [_P(c_void), c_int, c_void, c_void, c_void, c_void, c_void, c_void, c_void, c_int,
                                  c_void, c_int]),
    "grief_plan_destroy": (None, [c_void]),
    "grief_plan_info": (c_int, [c_void, c_int]),
    "grief_table_rows": (c_i64, [c_i64]),
    "grief_build_tables": (c_int, [c_void, c_void, c_i64, c_i64, c_void, c_void]),
    "grief_build_tables_dx": (c_int, [c_void, c_void, c_i64, c_i64, c_int, c_void, c_void]),
    "grief_build_tables_kxu": (c_int, [c_void, c_void, c_i64, c_void, c_i64, c_i64, c_int, c_void, c_void]),
    "grief_phi_rows": (c_int, [c_void, c_void, c_i64, c_void, c_void]),
    "grief_gram_workspace_bytes": (c_size, [c_void, c_i64]),
    "grief_gram": (c_int, [c_void, c_void, c_i64, c_void, c_i64, c_void, c_size, c_void]),
    "grief_phi_t_vec_workspace_bytes": (c_size, [c_void, c_i64]),
    "grief_phi_t_vec": (c_int, [c_void, c_void, c_i64, c_void, c_void, c_void, c_void]),
    "grief_phi_vec": (c_int, [c_void, c_void, c_i64, c_void, c_void, c_void]),
    "grief_sumsq": (c_int, [c_void, c_i64, c_void, c_void, c_void]),
    "grief_grad_setup": (c_int, [c_void, c_int, c_void, c_void, c_void]),
    "grief_grad_workspace_bytes": (c_size, [c_void, c_i64]),
    "grief_grad_theta": (c_int, [c_void, c_void, c_void, c_i64, c_void, c_i64, c_void, c_i64, c_void, c_dbl, c_void, c_void,
                                 c_void, c_size, c_void]),
    "grief_quadform_workspace_bytes": (c_size, [c_void, c_i64]),
    "grief_quadform_rows": (c_int, [c_void, c_void, c_i64, c_void, c_i64, c_void, c_void, c_size, c_void]),
    "grief_gram_ry": (c_int, [c_void, c_void, c_i64, c_void, c_void, c_i64, c_void, c_void, c_void, c_size, c_void]),
    "grief_set_default_option": (c_int, [c_int, c_i64]),
    "grief_get_default_option": (c_i64, [c_int]),
    "grief_plan_set_option": (c_int, [c_void, c_int, c_i64]),
    "grief_plan_get_option": (c_i64, [c_void, c_int]),
    "grief_set_slab_budget": (None, [c_size]),
    "grief_set_gemm_mode": (None, [c_int]),
    "grief_get_gemm_mode": (c_int, []),
    "grief_comm_unique_id": (c_int, [c_void]),
    "grief_comm_create": (c_int, [c_void, c_void, c_int, c_int]),
    "grief_comm_allreduce_sum": (c_int, [c_void, c_void, c_i64, c_void]),
    "grief_comm_destroy": (None, [c_void]),
    "grief_rowcol_kr_matvec": (c_int, [c_int, c_void, c_void, c_void, c_void, c_i64, c_i64, c_void, c_void, c_void]),
    "grief_gemm_nt": (c_int, [c_void, c_i64, c_void, c_i64, c_void, c_i64, c_int, c_int, c_int, c_dbl, c_dbl, c_void]),
    "grief_gemm_nt_t": (c_int, [c_void, c_i64, c_void, c_i64, c_void, c_i64, c_int, c_int, c_int, c_dbl, c_dbl, c_void]),
    "grief_solve_lml": (c_int, [c_void, c_int, c_void, c_i64, c_void, c_void, c_void, c_dbl, c_i64, c_void, c_void,
                                c_void, c_void, c_void, c_void, _P(c_int), c_void]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """The loaded library (loads on first use; raises NativeLibraryError if it cannot)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(gp_grief_b200/csrc/build.sh). There is no CPU fallback." % LIB_PATH)
        try:
            handle = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        except OSError as e:  # pragma: no cover
            raise NativeLibraryError("cannot load %s: %s" % (LIB_PATH, e))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    msg = lib().grief_last_error()
    return msg.decode() if msg else ""


def check(rc):
    """Map a C return code to the exception the reference would have raised."""
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_NOT_PD:
        raise np.linalg.LinAlgError(msg)
    if rc == ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError("libgrief_b200 error %d: %s" % (rc, msg))


PROFILE_SLOTS = ("k_gram", "k_zgemm", "k_tables", "k_contract", "k_topk", "solve", "phi_t_y", "k_dtables", "k_build_phi_t",
                 "k_build_phi")


def profile_enable(on=True):
    lib().grief_profile_enable(1 if on else 0)


def profile_read():
    """{slot name: (milliseconds, launches)} accumulated since the last read."""
    n = lib().grief_profile_slots()
    ms = np.zeros(n, dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int32)
    lib().grief_profile_read(host_ptr(ms), host_ptr(cnt))
    return {PROFILE_SLOTS[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}


def host_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def dev_ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
