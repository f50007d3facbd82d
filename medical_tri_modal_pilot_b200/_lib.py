"""ctypes binding of libtmp_b200.so (C ABI in include/tmp_b200.h).

The product path has NO CPU / PyTorch fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TMP_B200_LIB") or os.path.join(_PKG, "libtmp_b200.so")   # env override: A/B debugging only

_vp, _i, _ll, _f, _u32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint32
_pp = C.POINTER(C.c_void_p)

# name -> argtypes (restype is always int unless listed in _RESTYPES)
SIGNATURES = {
    "tmp_abi_version": [],
    "tmp_last_error": [],
    "tmp_set_reserved_sms": [_i],
    "tmp_num_sms": [],
    "tmp_build_lengths": [_vp, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "tmp_debug_materialize_mask": [_vp, _i, _i, _vp, _vp],
    "tmp_umse_embed_fwd": [_vp, _ll, _pp, _pp, _vp, _vp, _i, _vp],
    "tmp_stream_prologue_fwd": [_i, _i, _i, _vp, _pp, _vp, _vp, _i, _i, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _u32,
                                _u32, _vp, _vp, _vp],
    "tmp_stream_prologue_bwd": [_i, _i, _i, _vp, _pp, _vp, _vp, _i, _i, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _u32,
                                _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "tmp_layernorm_fwd": [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp],
    "tmp_layernorm_bwd": [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _f, _u32, _u32, _vp, _vp, _vp, _vp],
    "tmp_layernorm_bwd_attn": [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp],
    "tmp_gemm_bias_act_fwd": [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _f, _vp, _i, _vp, _i, _i, _vp, _i, _i, _f, _u32,
                              _u32, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp],
    "tmp_gemm_wgrad": [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "tmp_colsum": [_vp, _i, _ll, _i, _vp, _vp],
    "tmp_mma_attn_fwd": [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _i, _i, _vp],
    "tmp_attn_bwd_single_query": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "tmp_mma_attn_bwd": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp],
    "tmp_bottleneck_mix_fwd": [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    "tmp_bottleneck_mix_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _f, _u32, _vp, _u32, _u32, _u32, _vp],
    "tmp_dropout_apply": [_vp, _vp, _ll, _f, _u32, _u32, _vp, _vp],
    "tmp_cast_weights": [_vp, _i, _i, _i, _vp],
    "tmp_adamw_step": [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _vp],
    "tmp_adamw_step_dev": [_vp, _vp, _vp, _vp, _ll, _vp, _f, _f, _f, _f, _vp, _i, _vp],
    "tmp_grad_nonfinite": [_vp, _ll, _vp, _vp],
    "tmp_head_fwd": [_vp, _vp, _vp, _i, _pp, _vp, _vp, _vp, _f, _f, _pp, _vp, _vp, _vp, _vp],
    "tmp_head_bwd": [_vp, _vp, _vp, _i, _pp, _pp, _pp, _vp, _vp, _vp, _vp, _vp],
    # fp32 ("precise") mode: same operators on fp32-stored tensors, bf16x3 operand split, CUDA-core fp32 attention
    "tmp_layernorm_fwd_f32": [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp],
    "tmp_layernorm_bwd_f32": [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _f, _u32, _u32, _vp, _vp, _vp, _vp],
    "tmp_bottleneck_mix_fwd_f32": [_vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp],
    "tmp_bottleneck_mix_bwd_f32": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _f, _u32, _vp, _u32, _u32, _u32,
                                   _vp],
    "tmp_dropout_apply_f32": [_vp, _vp, _ll, _f, _u32, _u32, _vp, _vp],
    "tmp_colsum_f32": [_vp, _i, _ll, _i, _vp, _vp],
    "tmp_stream_prologue_fwd_f32": [_i, _i, _i, _vp, _pp, _vp, _vp, _i, _i, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _u32,
                                    _u32, _vp, _vp, _vp],
    "tmp_stream_prologue_bwd_f32": [_i, _i, _i, _vp, _pp, _vp, _vp, _i, _i, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _u32,
                                    _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "tmp_split_bf16x3": [_vp, _ll, _ll, _i, _i, _i, _vp, _vp],
    "tmp_attn_fwd_f32": [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _i, _vp],
    "tmp_attn_bwd_f32": [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp],
    "tmp_swin_patch_embed_ln": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp],
    "tmp_swin_ln_window": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "tmp_swin_window_attn": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp],
    "tmp_swin_unwindow_add_ln": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "tmp_swin_merge_ln": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
}
_RESTYPES = {"tmp_last_error": C.c_char_p}

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library (built by medical_tri_modal_pilot_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -m medical_tri_modal_pilot_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback for this path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def last_error() -> str:
    return (load().tmp_last_error() or b"").decode()


# kernels launched per C-ABI call (tmp_mma_attn_bwd = delta + main + dQ convert); bench.py reports the total
_KERNELS_PER_CALL = {"tmp_mma_attn_bwd(standalone)": 3, "tmp_attn_bwd_f32": 2, "tmp_head_bwd": 2}
launch_count = 0


def check(rc: int, what: str) -> None:
    global launch_count
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")
    launch_count += _KERNELS_PER_CALL.get(what, 1)


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr_array(tensors) -> "C.Array":
    """HOST array of device pointers (for the `const float* const*` branch-parameter blocks)."""
    arr = (C.c_void_p * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr
