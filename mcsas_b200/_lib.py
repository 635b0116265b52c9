# -*- coding: utf-8 -*-
"""ctypes binding of libmcsas_b200.so (include/mcsas_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
usable, the product path raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmcsas_b200.so")

MAX_PARAMS, MAX_ACTIVE, MAX_NQ, MAX_SPEC = 16, 4, 4096, 32
FP64, FP32 = 0, 1
QUAD_REFERENCE, QUAD_GAUSS_LEGENDRE = 0, 1
GEN_UNIFORM, GEN_EXP1, GEN_EXP2, GEN_EXP3 = 0, 1, 2, 3

EXPORTS = [
    "mcsas_ctx_create", "mcsas_ctx_destroy", "mcsas_last_error", "mcsas_device_info",
    "mcsas_upload_datasets", "mcsas_set_model", "mcsas_set_smearing", "mcsas_eval", "mcsas_run", "mcsas_run_async",
    "mcsas_wait", "mcsas_request_stop", "mcsas_run_info", "mcsas_scheduler_info", "mcsas_fetch", "mcsas_fetch_trace", "mcsas_fractions",
    "mcsas_measure_peaks", "mcsas_set_tuning", "mcsas_prepare_datasets",
]


class McSASB200Error(RuntimeError):
    pass


class RunConfig(C.Structure):
    """struct mcsas_run_config"""
    _fields_ = [("n_contribs", C.c_int), ("n_reps", C.c_int), ("conv_crit", C.c_double),
                ("max_iter", C.c_longlong), ("max_retries", C.c_int), ("comp_exp", C.c_double),
                ("find_background", C.c_int), ("positive_background", C.c_int),
                ("start_from_minimum", C.c_int), ("precision", C.c_int),
                ("seed", C.c_ulonglong), ("spec_max", C.c_int), ("record_trace", C.c_int),
                ("first_chain", C.c_int), ("cluster_size", C.c_int), ("slice_steps", C.c_longlong)]


class Tuning(C.Structure):
    """struct mcsas_tuning (experiment knobs, all 0 = defaults)"""
    _fields_ = [("max_cluster", C.c_int), ("slice_us", C.c_int), ("backups", C.c_int),
                ("backup_div", C.c_int), ("gram", C.c_int), ("qpt", C.c_int), ("profile", C.c_int),
                ("verbose", C.c_int), ("half", C.c_int), ("backup_pct", C.c_int), ("hedge_pct", C.c_int), ("early_pct", C.c_int), ("early_top", C.c_int), ("late_steps", C.c_int), ("big_batch", C.c_int), ("gram_in", C.c_int), ("gram_out", C.c_int), ("gpc_pack", C.c_int),
                ("l2_keep_pct", C.c_int), ("backup_top", C.c_int), ("retry_pct", C.c_int), ("repack_pct", C.c_int)]


_lib = None


def lib():
    """Load the shared library (built in-tree by __graft_entry__.build / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McSASB200Error(
            "libmcsas_b200.so not found at %s: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or make -C mcsas_b200/csrc). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    llp = C.POINTER(C.c_longlong)
    vp = C.c_void_p
    L.mcsas_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.mcsas_ctx_destroy.argtypes = [vp]
    L.mcsas_ctx_destroy.restype = None
    L.mcsas_last_error.argtypes = [vp]
    L.mcsas_last_error.restype = C.c_char_p
    L.mcsas_device_info.argtypes = [vp, ip, ip, C.c_char_p, C.c_int]
    L.mcsas_upload_datasets.argtypes = [vp, C.c_int, ip, dp, dp, dp]
    L.mcsas_set_model.argtypes = [vp, C.c_int, dp, C.c_int, ip, C.c_int, dp, dp, ip, C.c_int, C.c_int]
    L.mcsas_set_smearing.argtypes = [vp, C.c_int, dp, dp, C.c_int]
    L.mcsas_eval.argtypes = [vp, C.c_int, dp, C.c_int, C.c_double, C.c_int, dp, dp, dp, dp]
    L.mcsas_run.argtypes = [vp, C.POINTER(RunConfig), dp, C.c_longlong, ip]
    L.mcsas_run_async.argtypes = [vp, C.POINTER(RunConfig), dp, C.c_longlong, ip]
    L.mcsas_wait.argtypes = [vp, C.POINTER(C.c_float)]
    L.mcsas_request_stop.argtypes = [vp]
    L.mcsas_run_info.argtypes = [vp, ip, ip, C.POINTER(C.c_float)]
    L.mcsas_scheduler_info.argtypes = [vp, ip, ip]
    L.mcsas_fetch.argtypes = [vp, dp, dp, dp, dp, dp, dp, llp, llp, ip, dp, llp, llp]
    L.mcsas_fetch_trace.argtypes = [vp, C.c_int, C.POINTER(C.c_ubyte), dp, C.c_longlong]
    L.mcsas_fractions.argtypes = [vp, C.POINTER(RunConfig)] + [dp] * 11
    L.mcsas_measure_peaks.argtypes = [vp, dp, dp]
    L.mcsas_set_tuning.argtypes = [vp, C.POINTER(Tuning)]
    L.mcsas_prepare_datasets.argtypes = [vp, C.c_int, ip, dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                         C.c_int, C.c_int, C.c_int, dp, C.c_int, ip, dp, dp, dp]
    _lib = L
    return L
