"""ctypes binding of the C ABI in ``include/d2r_b200.h`` (``csrc/libd2r_b200.so``).

This is the only place that touches the shared library.  There is no fallback of any kind:
if the library is missing the import raises, and every non-zero status from the C side is
turned into a ``RuntimeError`` carrying ``d2r_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# D2R_B200_LIB: bring-up / bisection only -- load another in-tree build of the same sources
LIB_PATH = os.environ.get("D2R_B200_LIB") or os.path.join(_HERE, "csrc", "libd2r_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
EPI_STD, EPI_SQDIFF, EPI_SOFTMAX, EPI_SOFTMAX_BWD = 0, 1, 2, 3
ACT = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "tanh": ACT_TANH}


class GemmArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("c_dtype", C.c_int32),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("batch", C.c_int32), ("batch_inner", C.c_int32),
        ("act", C.c_int32), ("epilogue", C.c_int32), ("r_dtype", C.c_int32),
        ("accumulate", C.c_int32), ("split_k", C.c_int32),
        ("alpha", C.c_float), ("tile_n", C.c_int32), ("act_cols", C.c_int32), ("reserved0", C.c_int32),
        ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p), ("c2", C.c_void_p),
        ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64), ("ldr", C.c_int64),
        ("a_so", C.c_int64), ("a_si", C.c_int64), ("b_so", C.c_int64), ("b_si", C.c_int64),
        ("c_so", C.c_int64), ("c_si", C.c_int64), ("r_so", C.c_int64), ("r_si", C.c_int64),
        ("bias_sz", C.c_int64),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("heads", C.c_int32), ("Lq", C.c_int32), ("Lc", C.c_int32), ("hd", C.c_int32),
        ("mode", C.c_int32), ("alpha", C.c_float), ("reserved0", C.c_int32),
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("q_ld", C.c_int64), ("k_ld", C.c_int64), ("v_ld", C.c_int64),
        ("p", C.c_void_p), ("p_ld", C.c_int64),
        ("out", C.c_void_p), ("out2", C.c_void_p), ("o_ld", C.c_int64),
        ("residual", C.c_void_p), ("r_ld", C.c_int64),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("heads", C.c_int32), ("Lq", C.c_int32), ("Lc", C.c_int32), ("hd", C.c_int32),
        ("reserved0", C.c_int32), ("alpha", C.c_float), ("sign", C.c_float),
        ("d_out", C.c_void_p), ("p", C.c_void_p), ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("do_ld", C.c_int64), ("p_ld", C.c_int64), ("q_ld", C.c_int64), ("k_ld", C.c_int64), ("v_ld", C.c_int64),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("dq_ld", C.c_int64), ("dk_ld", C.c_int64), ("dv_ld", C.c_int64),
    ]


class Ptr8(C.Structure):
    _fields_ = [("p", C.c_void_p * 8)]


class AggArgs(C.Structure):
    _fields_ = [
        ("K", C.c_int32), ("n_out", C.c_int32), ("final_layer", C.c_int32), ("dtype", C.c_int32),
        ("B", C.c_int64), ("L", C.c_int64), ("D", C.c_int64),
        ("full", Ptr8), ("bvec", Ptr8), ("inputs", Ptr8), ("out", Ptr8),
        ("P", C.c_void_p), ("gate", C.c_void_p), ("pooled", C.c_void_p),
    ]


class AggBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", AggArgs), ("d_out", Ptr8), ("d_pooled", C.c_void_p),
        ("d_full", Ptr8), ("d_bvec", Ptr8), ("dP", C.c_void_p),
    ]


class SafArgs(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("training", C.c_int32),
        ("B", C.c_int64), ("L", C.c_int64), ("D", C.c_int64),
        ("sg", C.c_void_p), ("sl", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p),
        ("bn_w", C.c_void_p), ("bn_b", C.c_void_p),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
        ("logits", C.c_void_p), ("attn", C.c_void_p), ("stats", C.c_void_p), ("rnorm", C.c_void_p),
        ("out", C.c_void_p),
    ]


class SafBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", SafArgs), ("d_out", C.c_void_p), ("d_sg", C.c_void_p), ("d_sl", C.c_void_p),
        ("d_w", C.c_void_p), ("d_bias", C.c_void_p), ("d_bn_w", C.c_void_p), ("d_bn_b", C.c_void_p),
        ("scratch", C.c_void_p),
    ]


# every symbol include/d2r_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "d2r_abi_version": (C.c_int, []),
    "d2r_build_arch": (C.c_char_p, []),
    "d2r_last_error": (C.c_char_p, []),
    "d2r_launch_count": (C.c_int64, []),
    "d2r_gemm": (C.c_int, [C.POINTER(GemmArgs), _vp]),
    "d2r_gemm_set_profile": (C.c_int, [_vp]),
    "d2r_attn_fwd": (C.c_int, [C.POINTER(AttnArgs), _vp]),
    "d2r_attn_bwd": (C.c_int, [C.POINTER(AttnBwdArgs), _vp]),
    "d2r_softmax_fwd": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i32, _f, _vp]),
    "d2r_softmax_bwd": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _vp, _i32, _i64, _i64, _i32, _f, _vp]),
    "d2r_pool_mean": (C.c_int, [Ptr8, _i32, _i32, _i64, _i64, _i64, _vp, _vp]),
    "d2r_pool_mean_bwd": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _i32, _i32, _vp]),
    "d2r_router_head_fwd": (C.c_int, [_vp, Ptr8, Ptr8, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "d2r_router_head_bwd": (C.c_int, [_vp, _vp, _vp, Ptr8, _i32, _i32, _i64, _i32, _i32, _vp, _vp, Ptr8, Ptr8, _vp]),
    "d2r_aggregate_fwd": (C.c_int, [C.POINTER(AggArgs), _vp]),
    "d2r_aggregate_bwd": (C.c_int, [C.POINTER(AggBwdArgs), _vp]),
    "d2r_gate_skip_bwd": (C.c_int, [_vp, _vp, _vp, Ptr8, _i32, _i64, _i64, _i64, _i32, _i32, _vp]),
    "d2r_cast": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _vp]),
    "d2r_bias_act_bwd": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _i64, _i32, _i64, _vp]),
    "d2r_l2norm_fwd": (C.c_int, [_vp, _i32, _vp, _vp, _i64, _i32, _vp]),
    "d2r_l2norm_bwd": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp]),
    "d2r_film_fwd": (C.c_int, [_vp, _vp, _i32, _vp, _i64, _i32, _vp]),
    "d2r_film_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp]),
    "d2r_axpby": (C.c_int, [_vp, _vp, _i32, _f, _f, _vp, _i64, _vp]),
    "d2r_sqdiff_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp]),
    "d2r_mul": (C.c_int, [_vp, _vp, _i32, _f, _vp, _i64, _vp]),
    "d2r_saf_fwd": (C.c_int, [C.POINTER(SafArgs), _vp]),
    "d2r_saf_bwd": (C.c_int, [C.POINTER(SafBwdArgs), _vp]),
    "d2r_gate_fuse_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "d2r_gate_fuse_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "d2r_js_div_fwd": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "d2r_js_div_bwd": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "d2r_block_merge_fwd": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "d2r_block_merge_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"d2r_b200: CUDA library not built ({LIB_PATH} missing). Run `python -m d2r_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.d2r_abi_version() != 1:
        raise ImportError("d2r_b200: ABI version mismatch between _lib.py and libd2r_b200.so")
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.d2r_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"d2r_b200 {what} failed (status {rc}): {msg}")


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"d2r_b200: unsupported dtype {t.dtype}")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    """The current stream of the CURRENT device: callers run under ``torch.cuda.device(operand device)``
    (autograd.py does this for every stack / cell call) and require_cuda() rejects operands of another device."""
    return torch.cuda.current_stream().cuda_stream


def ptr8(tensors) -> Ptr8:
    p = Ptr8()
    for i, t in enumerate(tensors):
        p.p[i] = None if t is None else t.data_ptr()
    return p


def launch_count() -> int:
    return int(lib.d2r_launch_count())


def require_cuda(*tensors) -> None:
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("d2r_b200: tensors must live on a CUDA device (there is no CPU path)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"d2r_b200: operand on cuda:{t.device.index} but the current device is cuda:{cur}; "
                               "wrap the call in torch.cuda.device(tensor.device) (kernels are launched on the "
                               "current device's stream)")
