"""ctypes binding of libnisb200.so — the C-ABI CUDA library (include/nis_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is present, the entry points
raise ``NisBackendError``.
"""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NIS_LIB_PATH") or os.path.join(HERE, "libnisb200.so")     # override: development A/B builds

NIS_MAX_DIM = 32
NIS_MAX_CELLS = 32
NIS_MAX_HIDDEN = 8
NIS_MAX_WIDTH = 512
NIS_MAX_FINAL = 8
KIND_PWLIN, KIND_PWQUAD, KIND_AFFINE = 0, 1, 2
F32, F64 = 0, 1
BN_EVAL, BN_TRAIN = 0, 1


class NisBackendError(RuntimeError):
    pass


class NisCellDesc(ctypes.Structure):
    _fields_ = [("n_pass", ctypes.c_int32),
                ("feed_idx", ctypes.c_int32 * NIS_MAX_DIM),
                ("trafo_idx", ctypes.c_int32 * NIS_MAX_DIM),
                ("param_off", ctypes.c_int64),
                ("bn_off", ctypes.c_int64)]


class NisFlowDesc(ctypes.Structure):
    _fields_ = [("n_flow", ctypes.c_int32), ("n_cells", ctypes.c_int32), ("kind", ctypes.c_int32),
                ("n_bins", ctypes.c_int32), ("depth", ctypes.c_int32),
                ("widths", ctypes.c_int32 * NIS_MAX_HIDDEN),
                ("out_perm", ctypes.c_int32 * NIS_MAX_DIM),
                ("bn_eps", ctypes.c_float), ("bn_momentum", ctypes.c_float),
                ("cells", NisCellDesc * NIS_MAX_CELLS)]


class NisRamboDesc(ctypes.Structure):
    _fields_ = [("n_final", ctypes.c_int32),
                ("initial_masses", ctypes.c_double * 2),
                ("final_masses", ctypes.c_double * NIS_MAX_FINAL),
                ("E_cm", ctypes.c_double), ("pT_mincut", ctypes.c_double),
                ("delR_mincut", ctypes.c_double), ("rap_maxcut", ctypes.c_double),
                ("pdf_active", ctypes.c_int32), ("tau_mode", ctypes.c_int32),
                ("tau_min", ctypes.c_double), ("x_cut", ctypes.c_double),
                ("pdf_grid", ctypes.c_void_p * 2), ("pdf_nodes", ctypes.c_int32), ("pdf_lnx_lo", ctypes.c_double)]


_P = ctypes.c_void_p
_PROTOS = {
    "nis_flow_cell_param_count": (ctypes.c_int64, [ctypes.POINTER(NisFlowDesc), ctypes.c_int32]),
    "nis_flow_cell_bn_count": (ctypes.c_int64, [ctypes.POINTER(NisFlowDesc), ctypes.c_int32]),
    "nis_flow_bn_saved_count": (ctypes.c_int64, [ctypes.POINTER(NisFlowDesc)]),
    "nis_flow_workspace_bytes": (ctypes.c_size_t, [ctypes.POINTER(NisFlowDesc), ctypes.c_int64]),
    "nis_flow_forward": (ctypes.c_int, [ctypes.POINTER(NisFlowDesc), _P, _P, _P, ctypes.c_int32, ctypes.c_int32,
                                        _P, ctypes.c_int32, _P, _P, _P, ctypes.c_int32, _P, ctypes.c_size_t,
                                        ctypes.c_int64, _P]),
    "nis_flow_act_saved_count": (ctypes.c_int64, [ctypes.POINTER(NisFlowDesc), ctypes.c_int64]),
    "nis_flow_forward_cached": (ctypes.c_int, [ctypes.POINTER(NisFlowDesc), _P, _P, _P, ctypes.c_int32, ctypes.c_int32,
                                               _P, ctypes.c_int32, _P, _P, _P, _P, ctypes.c_int32, _P, ctypes.c_size_t,
                                               ctypes.c_int64, _P]),
    "nis_flow_backward_cached": (ctypes.c_int, [ctypes.POINTER(NisFlowDesc), _P, _P, _P, _P, _P, _P, ctypes.c_int32, _P, _P,
                                                ctypes.c_int32, _P, ctypes.c_size_t, ctypes.c_int64, _P]),
    "nis_flow_inverse": (ctypes.c_int, [ctypes.POINTER(NisFlowDesc), _P, _P, _P, ctypes.c_int32, ctypes.c_int32,
                                        _P, ctypes.c_int32, _P, ctypes.c_int32, _P, ctypes.c_size_t, ctypes.c_int64, _P]),
    "nis_flow_backward": (ctypes.c_int, [ctypes.POINTER(NisFlowDesc), _P, _P, _P, _P, _P, ctypes.c_int32, _P, _P,
                                         ctypes.c_int32, _P, ctypes.c_size_t, ctypes.c_int64, _P]),
    "nis_reduce_workspace_bytes": (ctypes.c_size_t, []),
    "nis_reduce_moments": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int64, _P, ctypes.c_int32, _P,
                                          ctypes.c_size_t, _P]),
    "nis_reduce_stats": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int64, _P, ctypes.c_int32, _P,
                                        ctypes.c_size_t, _P]),
    "nis_rambo_generate": (ctypes.c_int, [ctypes.POINTER(NisRamboDesc), _P, ctypes.c_int32, _P, _P, _P,
                                          ctypes.c_int64, _P]),
    "nis_rambo_invert": (ctypes.c_int, [ctypes.POINTER(NisRamboDesc), _P, _P, _P, ctypes.c_int64, _P]),
    "nis_uniform_fill": (ctypes.c_int, [_P, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint64, _P]),
    "nis_probe_fp32_fma": (ctypes.c_int64, [_P, ctypes.c_int32, _P]),
    "nis_probe_tensor": (ctypes.c_int64, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _P]),
    "nis_flow_timing_begin": (ctypes.c_int, [_P]),
    "nis_flow_timing_end": (ctypes.c_int, [_P, _P, ctypes.c_int32]),
    "nis_sizeof_flow_desc": (ctypes.c_size_t, []),
    "nis_sizeof_rambo_desc": (ctypes.c_size_t, []),
    "nis_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "nis_version": (ctypes.c_char_p, []),
}
EXPORTS = tuple(_PROTOS)

_lib = None


def load():
    """dlopen libnisb200.so and bind every symbol of include/nis_b200.h (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NisBackendError(
            "libnisb200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`python nf_b200/build.py`); nf_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.nis_sizeof_flow_desc() != ctypes.sizeof(NisFlowDesc) or \
            lib.nis_sizeof_rambo_desc() != ctypes.sizeof(NisRamboDesc):
        raise NisBackendError("ABI mismatch between nf_b200/_cabi.py and libnisb200.so")
    _lib = lib
    return lib


def lib():
    """The loaded library, after checking that a CUDA device exists (product entry points use this)."""
    if not torch.cuda.is_available():
        raise NisBackendError("nf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return load()


def check(rc, what):
    if rc != 0:
        raise NisBackendError("%s failed: %s (%d)" % (what, load().nis_strerror(rc).decode(), rc))


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise TypeError("nf_b200 kernels take float32 or float64 tensors, got %s" % t.dtype)
