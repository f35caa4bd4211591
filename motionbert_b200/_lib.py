"""ctypes binding of libmotionbert_b200.so (the C ABI in include/motionbert_b200.h).

This is the stub a maintainer of the reference would add.  It fails loudly
when the CUDA library is missing: there is no CPU or PyTorch fallback behind it.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MB_LIB_OVERRIDE") or os.path.join(HERE, "libmotionbert_b200.so")   # override: A/B builds only
TEST_LIB_PATH = os.path.join(HERE, "libmotionbert_b200_test.so")     # tests only: + reference kernels and hooks

MB_MATH_BF16X3 = 0
MB_MATH_BF16 = 1
MB_MATH_F16C = 2
MB_FLAG_REF_GEMM = 0x1
MB_FLAG_REF_ATTN_T = 0x2
MB_FLAG_GEMM_1CTA = 0x4
MB_FLAG_REF_ATTN_S = 0x8
MB_FLAG_ATTN_T_UNPACKED = 0x20
MB_FLAG_ATTN_BF16X3 = 0x40
MB_FLAG_MLP_SPLIT = 0x100      # F16C: the MLP sublayer as two GEMM launches instead of the fused kernel (A/B, tests)
MB_FLAG_MLP_NO_HINT = 0x800    # fused MLP: no L2::evict_last on the hidden stores / loads
MB_FLAG_MLP_NO_RING = 0x200    # fused MLP: hidden rows indexed by token block instead of the per-pair L2 ring

# every symbol include/motionbert_b200.h declares (TEST_EXPORTS: include/motionbert_b200_test.h, test library only)
EXPORTS = [
    "mb_version", "mb_last_error", "mb_create", "mb_destroy", "mb_param_count", "mb_param_info",
    "mb_packed_bytes", "mb_pack_weights", "mb_workspace_bytes", "mb_forward", "mb_workspace_bytes_host",
    "mb_forward_host", "mb_forward_pooled", "mb_forward_launch_count", "mb_profile_enable", "mb_profile_read",
    "mb_saved_bytes", "mb_forward_train", "mb_backward_workspace_bytes", "mb_backward", "mb_backward_launch_count", "mb_pretrain_loss", "mb_augment2d", "mb_adamw_step",
]
TEST_EXPORTS = [
    "mb_test_linear_scratch_bytes", "mb_test_linear",
    "mb_test_attention_scratch_bytes", "mb_test_attention", "mb_test_wgrad_scratch_bytes", "mb_test_wgrad",
    "mb_test_dgrad_scratch_bytes", "mb_test_dgrad",
    "mb_test_attention_backward_scratch_bytes", "mb_test_attention_backward", "mb_test_f16c_encode",
]


class MbDesc(C.Structure):
    _fields_ = [
        ("dim_in", C.c_int32), ("dim_out", C.c_int32), ("dim_feat", C.c_int32), ("dim_rep", C.c_int32),
        ("depth", C.c_int32), ("num_heads", C.c_int32), ("hidden", C.c_int32), ("num_joints", C.c_int32),
        ("maxlen", C.c_int32), ("eps", C.c_float), ("qk_scale", C.c_float), ("math", C.c_int32),
    ]


class MbError(RuntimeError):
    pass


_lib = None
_test_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree product library; raises if it has not been built (python -m motionbert_b200.build)."""
    global _lib
    if _lib is None:
        _lib = _open(LIB_PATH, test=False)
    return _lib


def load_test() -> C.CDLL:
    """dlopen the test twin (product ABI + reference kernels + kernel-level hooks).  tests/ only."""
    global _test_lib
    if _test_lib is None:
        _test_lib = _open(TEST_LIB_PATH, test=True)
    return _test_lib


def _open(path: str, test: bool) -> C.CDLL:
    if not os.path.exists(path):
        raise MbError(f"{path} is missing: build it with `python -m motionbert_b200.build` "
                      "(there is no CPU / PyTorch fallback for the DSTformer hot path)")
    lib = C.CDLL(path)
    vp, i32, u32, sz = C.c_void_p, C.c_int, C.c_uint32, C.c_size_t
    fp = C.c_void_p   # device / host float* passed as raw addresses
    lib.mb_version.restype = i32
    lib.mb_last_error.restype = C.c_char_p
    lib.mb_create.argtypes = [C.POINTER(MbDesc), C.POINTER(vp)]
    lib.mb_destroy.argtypes = [vp]
    lib.mb_destroy.restype = None
    lib.mb_param_count.argtypes = [vp]
    lib.mb_param_info.argtypes = [vp, i32, C.c_char_p, i32, C.POINTER(C.c_int64)]
    lib.mb_packed_bytes.argtypes = [vp, C.POINTER(sz)]
    lib.mb_pack_weights.argtypes = [vp, C.POINTER(vp), vp, vp]
    lib.mb_workspace_bytes.argtypes = [vp, i32, i32, C.POINTER(sz)]
    lib.mb_forward.argtypes = [vp, vp, fp, fp, fp, fp, vp, sz, i32, i32, u32, vp]
    lib.mb_workspace_bytes_host.argtypes = [vp, i32, i32, i32, i32, C.POINTER(sz)]
    lib.mb_forward_host.argtypes = [vp, vp, fp, fp, fp, vp, sz, i32, i32, u32, vp]
    lib.mb_forward_pooled.argtypes = [vp, vp, fp, fp, vp, sz, i32, i32, u32, vp]
    lib.mb_forward_launch_count.argtypes = [vp, i32, u32]
    lib.mb_profile_enable.argtypes = [vp, i32]
    lib.mb_profile_read.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.mb_saved_bytes.argtypes = [vp, i32, i32, C.POINTER(sz)]
    lib.mb_forward_train.argtypes = [vp, vp, fp, fp, fp, fp, vp, sz, vp, sz, i32, i32, u32, vp]
    lib.mb_backward_workspace_bytes.argtypes = [vp, i32, i32, C.POINTER(sz)]
    lib.mb_backward_launch_count.argtypes = [vp, i32, i32]
    lib.mb_backward.argtypes = [vp, vp, C.POINTER(vp), fp, fp, vp, sz, fp, fp, fp, C.POINTER(vp), fp, vp, sz, i32, i32, C.POINTER(vp), vp]
    lib.mb_pretrain_loss.argtypes = [fp, fp, fp, i32, i32, i32, C.c_float, C.c_float, fp, fp, vp, vp]
    f32 = C.c_float
    lib.mb_augment2d.argtypes = [fp, i32, i32, i32, i32, i32, i32, i32, fp, fp, fp, fp, fp, fp, fp, fp, f32, f32, f32, f32,
                                 f32, f32, fp, fp, f32, f32, fp, vp]
    lib.mb_adamw_step.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.c_char_p, i32, f32, f32, f32,
                                  f32, f32, vp]
    if test:
        _bind_hooks(lib)
    for name in EXPORTS + (TEST_EXPORTS if test else []):
        fn = getattr(lib, name)
        if name not in ("mb_last_error", "mb_destroy"):
            fn.restype = C.c_int
    return lib


def _bind_hooks(lib):
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    fp = C.c_void_p
    lib.mb_test_f16c_encode.argtypes = [fp, i32, i32, vp, vp]
    lib.mb_test_linear_scratch_bytes.argtypes = [i32, i32, i32, C.POINTER(sz)]
    lib.mb_test_linear.argtypes = [i32, i32, i32, i32, i32, i32, fp, fp, fp, fp, fp, fp, C.c_float, fp, fp, vp, sz, vp]
    lib.mb_test_attention_scratch_bytes.argtypes = [i32, i32, i32, i32, C.POINTER(sz)]
    lib.mb_test_attention.argtypes = [i32, i32, i32, i32, i32, i32, i32, i32, fp, fp, vp, sz, vp]
    lib.mb_test_wgrad_scratch_bytes.argtypes = [i32, i32, i32, C.POINTER(sz)]
    lib.mb_test_wgrad.argtypes = [i32, i32, i32, i32, fp, fp, fp, vp, sz, vp]
    lib.mb_test_dgrad_scratch_bytes.argtypes = [i32, i32, i32, C.POINTER(sz)]
    lib.mb_test_dgrad.argtypes = [i32, i32, i32, i32, fp, fp, fp, vp, sz, vp]
    lib.mb_test_attention_backward_scratch_bytes.argtypes = [i32, i32, i32, i32, C.POINTER(sz)]
    lib.mb_test_attention_backward.argtypes = [i32, i32, i32, i32, i32, i32, fp, fp, fp, vp, sz, vp]


def check(rc: int, what: str = "", lib: "C.CDLL | None" = None) -> int:
    """Raise MbError with the library's thread-local message on a negative status (`lib`: the library that returned it)."""
    if rc < 0:
        msg = (lib or load()).mb_last_error().decode("utf-8", "replace")
        raise MbError(f"{what or 'libmotionbert_b200'} failed ({rc}): {msg}")
    return rc
