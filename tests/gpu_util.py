"""Helpers for the -m gpu tests: call the C-ABI test hooks with torch-allocated device buffers."""
import ctypes

import torch

from motionbert_b200 import _lib


def _scratch(nbytes, device):
    t = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    return t, (t.data_ptr() + 1023) // 1024 * 1024


def test_linear(mode, A, W, bias, gamma=None, beta=None, resid=None, eps=1e-6, math=0, use_ref=0):
    """Returns (y fp32 [M,N], stats [M, N/128, 3] or None)."""
    lib = _lib.load_test()
    dev = A.device
    M, K = A.shape
    N = W.shape[0]
    nb = ctypes.c_size_t()
    _lib.check(lib.mb_test_linear_scratch_bytes(M, N, K, ctypes.byref(nb)), "hook", lib)
    keep, sp = _scratch(nb.value, dev)
    y = torch.full((M, N), float("nan"), dtype=torch.float32, device=dev)
    stats = torch.zeros(M, N // 128, 3, dtype=torch.float32, device=dev) if mode == 2 else None
    ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mb_test_linear(mode, math, use_ref, M, N, K, ptr(A), ptr(W), ptr(bias), ptr(gamma), ptr(beta),
                                      ptr(resid), eps, ptr(y), ptr(stats), sp, nb.value, st), "mb_test_linear", lib)
        torch.cuda.synchronize(dev)
    return y, stats


def test_attention(temporal, qkv, B, F, J, C, H, math=0, use_ref=0):
    lib = _lib.load_test()
    dev = qkv.device
    nb = ctypes.c_size_t()
    _lib.check(lib.mb_test_attention_scratch_bytes(B, F, J, C, ctypes.byref(nb)), "hook", lib)
    keep, sp = _scratch(nb.value, dev)
    y = torch.full((B * F * J, C), float("nan"), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mb_test_attention(int(temporal), math, use_ref, B, F, J, C, H, qkv.data_ptr(), y.data_ptr(),
                                         sp, nb.value, st), "mb_test_attention", lib)
        torch.cuda.synchronize(dev)
    return y


def test_wgrad(G_, X, math=0):
    """dW = G^T X through the tcgen05 weight-gradient kernel."""
    lib = _lib.load_test()
    dev = G_.device
    M, N = G_.shape
    K = X.shape[1]
    nb = ctypes.c_size_t()
    _lib.check(lib.mb_test_wgrad_scratch_bytes(M, N, K, ctypes.byref(nb)), "hook", lib)
    keep, sp = _scratch(nb.value, dev)
    dW = torch.full((N, K), float("nan"), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mb_test_wgrad(math, M, N, K, G_.data_ptr(), X.data_ptr(), dW.data_ptr(), sp, nb.value, st), "mb_test_wgrad", lib)
        torch.cuda.synchronize(dev)
    return dW


def test_dgrad(G_, W, math=0):
    """dX = G W through the 2-CTA GEMM with W consumed MN-major."""
    lib = _lib.load_test()
    dev = G_.device
    M, N = G_.shape
    K = W.shape[1]
    nb = ctypes.c_size_t()
    _lib.check(lib.mb_test_dgrad_scratch_bytes(M, N, K, ctypes.byref(nb)), "hook", lib)
    keep, sp = _scratch(nb.value, dev)
    dX = torch.full((M, K), float("nan"), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mb_test_dgrad(math, M, N, K, G_.data_ptr(), W.data_ptr(), dX.data_ptr(), sp, nb.value, st), "mb_test_dgrad", lib)
        torch.cuda.synchronize(dev)
    return dX


def test_attention_backward(temporal, qkv, dO, B, F, J, C, H):
    lib = _lib.load_test()
    dev = qkv.device
    nb = ctypes.c_size_t()
    _lib.check(lib.mb_test_attention_backward_scratch_bytes(B, F, J, C, ctypes.byref(nb)), "hook", lib)
    keep, sp = _scratch(nb.value, dev)
    dqkv = torch.full((B * F * J, 3 * C), float("nan"), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.mb_test_attention_backward(int(temporal), B, F, J, C, H, qkv.data_ptr(), dO.data_ptr(),
                                                  dqkv.data_ptr(), sp, nb.value, st), "mb_test_attention_backward", lib)
        torch.cuda.synchronize(dev)
    return dqkv


def f16c_encode(x):
    """The kernels' F16C encoder: fp32 [rows, cols] -> uint8 [rows, cols * 4]."""
    lib = _lib.load_test()
    rows, cols = x.shape
    out = torch.empty(rows, cols * 4, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.mb_test_f16c_encode(x.contiguous().data_ptr(), rows, cols, out.data_ptr(),
                                           torch.cuda.current_stream(x.device).cuda_stream), "mb_test_f16c_encode", lib)
        torch.cuda.synchronize(x.device)
    return out
