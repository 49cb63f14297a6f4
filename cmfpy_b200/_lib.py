"""ctypes binding of libcmf_b200.so (the C ABI declared in include/cmf_b200.h).

There is no CPU fallback: if the CUDA library is missing or no sm_100 GPU is
visible, solver construction raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMF_B200_LIB") or os.path.join(HERE, "lib", "libcmf_b200.so")   # (override: A/B builds)

CMF_F32, CMF_F64 = 0, 1
CMF_HOST, CMF_DEVICE = 0, 1
CMF_PREC_FP32, CMF_PREC_TF32, CMF_PREC_TF32X3 = 0, 1, 2
PRECISIONS = {"fp32": CMF_PREC_FP32, "tf32": CMF_PREC_TF32, "tf32x3": CMF_PREC_TF32X3}
CMF_DEN_DIRECT, CMF_DEN_GRAM, CMF_DEN_AUTO = 0, 1, 2
CMF_PEER_BLOB_BYTES = 512
DENOMINATORS = {"direct": CMF_DEN_DIRECT, "gram": CMF_DEN_GRAM, "auto": CMF_DEN_AUTO}
LOSS_MODES = {"auto": 0, "full": 1, "wterms": 2}           # cmf_mu_set_loss_mode


class SynthParams(C.Structure):
    _fields_ = [
        ("n_components", C.c_int), ("n_features", C.c_int), ("n_lags", C.c_int),
        ("n_timebins", C.c_longlong), ("t_offset", C.c_longlong), ("t_local", C.c_longlong),
        ("H_sparsity", C.c_double), ("noise_scale", C.c_double), ("seed", C.c_ulonglong),
        ("device", C.c_int), ("precision", C.c_int),
    ]


SYNTH_W, SYNTH_H, SYNTH_NOISE, SYNTH_DATA, SYNTH_GENERATE = range(5)


class Params(C.Structure):
    _fields_ = [
        ("n_features", C.c_int), ("n_components", C.c_int), ("maxlag", C.c_int),
        ("t_local", C.c_longlong), ("t_global", C.c_longlong), ("t_offset", C.c_longlong),
        ("device", C.c_int), ("precision", C.c_int), ("stream", C.c_void_p),
        ("denominators", C.c_int),
    ]


_H = C.c_void_p
_SIGNATURES = {
    "cmf_abi_version": (C.c_int, []),
    "cmf_last_error": (C.c_char_p, []),
    "cmf_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "cmf_release_cached_memory": (C.c_int, []),
    "cmf_precision_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "cmf_mu_create": (C.c_int, [C.POINTER(_H), C.POINTER(Params)]),
    "cmf_mu_destroy": (C.c_int, [_H]),
    "cmf_mu_set_data": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_longlong]),
    "cmf_mu_data_stats": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "cmf_mu_row_stats": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmf_mu_scale_rows": (C.c_int, [_H, C.c_void_p]),
    "cmf_mu_set_norm_x": (C.c_int, [_H, C.c_double]),
    "cmf_mu_set_factors": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong]),
    "cmf_mu_init_stats": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cmf_mu_scale_factors": (C.c_int, [_H, C.c_double, C.c_double]),
    "cmf_mu_halo_width": (C.c_int, [_H, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cmf_mu_halo_export": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "cmf_mu_halo_import": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "cmf_mu_recon": (C.c_int, [_H]),
    "cmf_mu_recon_loss": (C.c_int, [_H]),
    "cmf_mu_w_terms": (C.c_int, [_H]),
    "cmf_mu_w_terms_buffer": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]),
    "cmf_mu_w_apply": (C.c_int, [_H]),
    "cmf_mu_h_step": (C.c_int, [_H]),
    "cmf_mu_needs_mid_recon": (C.c_int, [_H, C.POINTER(C.c_int)]),
    "cmf_mu_resid_sumsq": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "cmf_mu_resid_sumsq_buffer": (C.c_int, [_H, C.POINTER(C.c_void_p)]),
    "cmf_mu_loss": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "cmf_mu_peer_attach_local": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(_H)]),
    "cmf_mu_halo_exchange_peer": (C.c_int, [_H]),
    "cmf_mu_step": (C.c_int, [_H, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "cmf_mu_peer_export": (C.c_int, [_H, C.c_void_p]),
    "cmf_mu_peer_attach": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p]),
    "cmf_mu_peer_detach": (C.c_int, [_H]),
    "cmf_mu_step_sharded": (C.c_int, [_H, C.c_int, C.POINTER(C.c_double)]),
    "cmf_hals_begin": (C.c_int, [_H]),
    "cmf_hals_sweep_w": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "cmf_hals_sweep_h": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "cmf_hals_end": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "cmf_gd_cache": (C.c_int, [_H]),
    "cmf_gd_lipschitz_w": (C.c_int, [_H, C.POINTER(C.c_double)]),
    "cmf_gd_lipschitz_state": (C.c_int, [_H, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cmf_gd_step": (C.c_int, [_H, C.c_int, C.c_double, C.POINTER(C.c_double)]),
    "cmf_mu_get_W": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int]),
    "cmf_mu_get_H": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.c_longlong]),
    "cmf_mu_get_est": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, C.c_longlong]),
    "cmf_mu_h_terms": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int]),
    "cmf_mu_get_w_terms": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int]),
    "cmf_mu_launch_count": (C.c_int, [_H, C.POINTER(C.c_longlong)]),
    "cmf_mu_path_name": (C.c_char_p, [_H]),
    "cmf_mu_kernel_ms": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "cmf_mu_set_profiling": (C.c_int, [_H, C.c_int]),
    "cmf_mu_set_loss_mode": (C.c_int, [_H, C.c_int]),
    "cmf_mu_launch_table": (C.c_int, [_H, C.c_char_p, C.c_longlong]),
    "cmf_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
                              C.c_int, C.c_int, C.c_int, C.c_int]),
    "cmf_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
                            C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "cmf_tensor_transconv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cmf_dmat_info": (C.c_int, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "cmf_dmat_get": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_longlong]),
    "cmf_dmat_destroy": (C.c_int, [_H]),
    "cmf_synth_create": (C.c_int, [C.POINTER(_H), C.POINTER(SynthParams)]),
    "cmf_synth_get": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int, C.c_longlong]),
    "cmf_synth_matrix": (C.c_int, [_H, C.c_int, C.POINTER(_H)]),
    "cmf_synth_destroy": (C.c_int, [_H]),
    "cmf_spectrogram": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_double, C.c_int, C.c_int,
                                  C.c_void_p, C.c_int, C.c_int, C.POINTER(_H)]),
}

_lib = None


def exported_symbols():
    """Names the header declares (used by the CPU-side ABI test)."""
    return sorted(_SIGNATURES)


def load():
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "cmfpy_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.cmf_abi_version() != 1:
        raise RuntimeError("cmfpy_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    """Turns a non-zero status into the Python exception the reference would
    raise (ValueError for argument errors, RuntimeError for CUDA failures)."""
    if rc == 0:
        return
    msg = load().cmf_last_error().decode("utf-8", "replace")
    if rc == 2:
        raise ValueError(msg)
    raise RuntimeError("cmfpy_b200: " + msg)


def launch_table(lib, handle):
    """{kernel label: (launches, total ms)} recorded since cmf_mu_set_profiling(handle, 2)."""
    buf = C.create_string_buffer(1 << 16)
    check(lib.cmf_mu_launch_table(handle, buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(" ", 2)
        out[name] = (int(n), float(ms))
    return out


def np_dtype_code(a):
    if a.dtype == np.float32:
        return CMF_F32
    if a.dtype == np.float64:
        return CMF_F64
    raise TypeError("expected float32 or float64, got %s" % a.dtype)


def release_cached_memory():
    """Returns the library's cached N x T device buffers (kept across solvers of the same shape) to the driver."""
    load().cmf_release_cached_memory()


def device_count():
    n = C.c_int(0)
    rc = load().cmf_device_count(C.byref(n))
    return n.value if rc == 0 else 0
