"""ctypes binding of libdrn_b200.so (the C ABI declared in include/drn_b200.h).

The product path has NO fallback: if the library is missing or a call fails, this raises.
PyTorch is only used by callers for device memory (`tensor.data_ptr()`) and the current stream.
"""
import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdrn_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "drn_b200.h")

DRN_F32, DRN_BF16, DRN_U8 = 0, 1, 2

_P = c_void_p
_FP = POINTER(c_float)
_IP = POINTER(c_int)

# name -> argtypes (restype is int unless noted)
_PROTOS = {
    "drn_version": [],
    "drn_conv3x3_c3_fwd": [_P, c_int, c_int, c_int, c_int, _FP, _FP, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P],
    "drn_conv_igemm_f32": [_P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P, c_int, _P, c_int, c_int, _P],
    "drn_conv_igemm_bf16_tc": [_P, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P, c_int, _P, c_int, c_int, c_int,
                               c_float, c_uint64, _P, _P, c_size_t, _P],
    "drn_maxpool2x2_nhwc": [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "drn_roipool_fwd": [_P, c_int, c_int, c_int, _P, _P, c_int, c_float, c_int, _P, _P, c_size_t, _P],
    "drn_roipool_tables_supported": [c_int, c_int],
    "drn_roipool_build_tables": [_P, c_int, c_int, c_int, c_int, _P, c_size_t, _P],
    "drn_roipool_rows_fwd": [_P, c_int, c_int, c_int, _P, _P, c_int, c_float, c_int, _P, _P, c_size_t, c_int, _P],
    "drn_wsddn_mil_fwd": [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_float, _P, _P, _P, _P, _P],
    "drn_oicr_pgt": [_P, c_int, c_int, _P, _P, c_int, _P, c_int, _P, c_int, c_int, _FP, _P, _P, _P, _P, _P],
    "drn_label_proposals": [_P, c_int, _P, _P, c_int, c_int, _FP, _IP, c_int, _P, _P, _P, _P],
    "drn_oicr_stage_fwd": [_P, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, c_float, _P, _P, _P, _P, _P, _P, _P],
    "drn_wsddn_mil_pgt_fwd": [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_float, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P,
                              _P, _P, _P],
    "drn_oicr_stage_fused_fwd": [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _FP, _IP, c_int, c_float, _P, _P, c_int,
                                 _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, _FP, _P, _P, _P, _P, _P, _P, _P],
    "drn_oicr_stages_fwd": [_P, c_int, c_int, c_int, c_int, _IP, _IP, _FP, _P, _P, c_int, _P, c_int, _P, _P, _FP, _IP, c_int, c_float,
                            _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _IP, _P, _P, c_int, _P],
    "drn_oicr_boxreg_loss": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _FP, c_float, c_float, _P, _P, _P, _P],
    "drn_oicr_infer": [_P, c_int, c_int, c_int, c_int, c_int, _IP, _IP, _P, _FP, _P, _P, _P],
    "drn_detections_fwd": [_P, _P, c_int, c_int, c_int, c_float, c_float, c_float, c_double, c_int, _P, _P, _P, _P, _P, _P, c_size_t,
                           _P],
    "drn_wsddn_mil_bwd": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, c_float, _P, _P, _P, _P],
    "drn_oicr_stage_bwd": [_P, _P, _P, _P, c_float, _P, c_int, c_int, c_int, c_int, _P, _P],
    "drn_oicr_boxreg_bwd": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _FP, c_float, c_float, c_float, _P, _P, _P],
    "drn_masked_transpose": [_P, c_int, c_int, _P, c_int, c_int, c_float, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, _P],
    "drn_rowsum": [_P, c_int, c_int, c_int, c_int, _P, _P],
    "drn_permute_cols49": [_P, _P, c_int64, c_int, _P],
    "drn_sgd_step": [_P, _P, _P, _P, c_int64, c_int64, c_int, c_float, c_float, c_float, c_int, c_int, _P],
    "drn_pack_linear_bf16": [_P, _P, c_int64, c_int64, c_int, _P],
    "drn_pcl_stage_fwd": [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int, c_float, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "drn_pcl_stage_bwd": [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, c_float, _P, c_int, c_int, _P, _P],
    "drn_peer_handle_bytes": [],
    "drn_peer_get_handle": [_P, _P, POINTER(c_uint64)],
    "drn_peer_open": [_P, POINTER(c_void_p)],
    "drn_peer_close": [_P],
    "drn_gemm_bf16_tc_scatter": [_P, c_int, c_int, _P, c_int, _P, POINTER(c_void_p), c_int, c_int, c_int, _P],
    "drn_sgd_step_sharded": [_P, _P, _P, c_int, POINTER(c_void_p), c_int, c_int64, c_int64, c_int64, c_int, c_float, c_float,
                             c_float, c_int, c_int, _P],
    "drn_dropout_inplace": [_P, c_int64, c_int, c_float, c_uint64, _P, _P],
    "drn_cast_f32_to_bf16": [_P, _P, c_int64, _P],
    "drn_cast_bf16_to_f32": [_P, _P, c_int64, _P],
    "drn_resample_u8_fwd": [_P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, c_int, c_int, c_int, _P, c_int, _P, c_int, _P],
    "drn_f32tc_split": [_P, c_int64, c_int, c_int, _P, _P, _P],
    "drn_f32tc_reduce": [_P, c_int, c_int64, c_int, _P, _P, c_int, c_int64, c_int, _P, _P],
    "drn_tta_accumulate": [_P, _P, c_int, c_int, c_int, c_int, _IP, _FP, _FP, _P, _P, c_int, c_int, _P],
}

_lib = None


def header_symbols():
    """Every function name declared in include/drn_b200.h."""
    with open(HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(drn_[a-z0-9_]+)\s*\(", src)))


def load():
    """Load the shared library (raises with a build hint if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is mandatory (no CPU/eager fallback). "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` from the repo root."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.drn_last_error.restype = c_char_p
    lib.drn_last_error.argtypes = []
    for name, argtypes in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.drn_roipool_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.drn_roipool_workspace_bytes.restype = c_size_t
    lib.drn_detections_workspace_bytes.argtypes = [c_int, c_int]
    lib.drn_detections_workspace_bytes.restype = c_size_t
    lib.drn_gemm_set_max_sms.argtypes = [c_int]
    lib.drn_gemm_set_max_sms.restype = c_int
    lib.drn_gemm_set_tail_split.argtypes = [c_int]
    lib.drn_gemm_set_tail_split.restype = c_int
    lib.drn_gemm_workspace_bytes.argtypes = []
    lib.drn_gemm_workspace_bytes.restype = c_size_t
    _lib = lib
    return lib


def _ptr(t):
    if t is None:
        return None
    return t.data_ptr()


def fvec(vals):
    return (c_float * len(vals))(*[float(v) for v in vals])


def ivec(vals):
    return (c_int * len(vals))(*[int(v) for v in vals])


def pvec(ptrs):
    """Host array of device pointers (void* const*)."""
    return (c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def call(name, *args):
    """Invoke a C-ABI entry point; tensors are passed as device pointers."""
    lib = load()
    conv = []
    for a in args:
        if hasattr(a, "data_ptr"):
            conv.append(a.data_ptr())
        else:
            conv.append(a)
    rc = getattr(lib, name)(*conv)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.drn_last_error().decode()}")
    return rc


def current_stream():
    import torch

    return torch.cuda.current_stream().cuda_stream
