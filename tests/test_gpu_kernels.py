"""Kernel-level parity through the C-ABI test hooks: each CUDA kernel against a plain torch fp64 restatement
of the same op (tolerances written per test).  Localises failures before the whole-model tests run."""
import math

import pytest
import torch

import gpu_util as G

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30)), float((a - b).abs().max())


def _mk(M, N, K, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = (torch.randn(M, K, generator=g) * 1.5 + 0.3).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    b = (torch.randn(N, generator=g) * 0.1).to(dev)
    gamma = (1 + 0.2 * torch.randn(K, generator=g)).to(dev)
    beta = (0.1 * torch.randn(K, generator=g)).to(dev)
    resid = torch.randn(M, N, generator=g).to(dev)
    return A, W, b, gamma, beta, resid


def _expected(mode, A, W, b, gamma, beta, resid, eps):
    A, W, b = A.double(), W.double(), b.double()
    if mode in (0, 1, 3):
        x = torch.nn.functional.layer_norm(A, (A.shape[1],), gamma.double(), beta.double(), eps)
        y = x @ W.T + b
        if mode == 1:
            y = torch.nn.functional.gelu(y)
        if mode == 3:
            y = torch.tanh(y)
        return y
    y = A @ W.T + b
    if mode == 2:
        y = resid.double() + y
    return y


SHAPES = [(128, 256, 256), (300, 512, 512), (1000, 1536, 512), (77, 512, 1024), (4131, 1024, 512), (459, 768, 256),
          (40000, 512, 512)]


@pytest.mark.parametrize("use_ref", [1, 2, 0], ids=["simt_ref", "tc1cta", "tc2cta"])
@pytest.mark.parametrize("mode", [4, 0, 1, 2, 3], ids=["bias", "ln_split", "ln_gelu", "resid", "ln_tanh"])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_bf16x3(cuda_device, M, N, K, mode, use_ref):
    A, W, b, gamma, beta, resid = _mk(M, N, K, cuda_device, seed=M + N + K + mode)
    y, stats = G.test_linear(mode, A, W, b, gamma, beta, resid, 1e-6, math=0, use_ref=use_ref)
    exp = _expected(mode, A, W, b, gamma, beta, resid, 1e-6)
    assert torch.isfinite(y).all(), "non-finite / unwritten output"
    rel, mx = _rel(y, exp)
    # BF16x3: 16-bit operand mantissas, fp32 accumulate; split-plane outputs (modes 0,1) carry 2^-17 rounding
    assert rel < 3e-5, f"rel {rel:.3e} max {mx:.3e}"
    if mode == 2:
        # LN partial statistics of the output rows: (shift, sum(x-shift), sum((x-shift)^2)) per 128 columns
        e = exp.float().reshape(M, N // 128, 128)
        mean_g = stats[..., 0] + stats[..., 1] / 128
        var_g = stats[..., 2] / 128 - (stats[..., 1] / 128) ** 2
        assert float((mean_g - e.mean(-1)).abs().max()) < 1e-4
        assert float((var_g - e.var(-1, unbiased=False)).abs().max()) < 1e-3


@pytest.mark.parametrize("mode", [4, 0, 1, 2, 3], ids=["bias", "ln_split", "ln_gelu", "resid", "ln_tanh"])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_f16c(cuda_device, M, N, K, mode):
    """F16C arithmetic (ptx.cuh): a w ~ ah wh (fp16 MMA) + q(al 2^6) q(wh 2^-6) + q(ah 2^-6) q(wl 2^6) (e5m2 MMAs, K = 32).
    Operands carry ~14-15 significant bits -> ~3e-5 per GEMM (BF16x3: ~1e-5); the split-row outputs (modes 0, 1) are
    decoded as h + l (same 14-15 bits)."""
    A, W, b, gamma, beta, resid = _mk(M, N, K, cuda_device, seed=M + N + K + mode)
    y, stats = G.test_linear(mode, A, W, b, gamma, beta, resid, 1e-6, math=2, use_ref=0)
    exp = _expected(mode, A, W, b, gamma, beta, resid, 1e-6)
    assert torch.isfinite(y).all(), "non-finite / unwritten output"
    rel, mx = _rel(y, exp)
    print(f"f16c M={M} N={N} K={K} mode={mode}: rel {rel:.3e} max {mx:.3e}")
    assert rel < 1.5e-4, f"rel {rel:.3e} max {mx:.3e}"
    if mode == 2:
        e = exp.float().reshape(M, N // 128, 128)
        mean_g = stats[..., 0] + stats[..., 1] / 128
        var_g = stats[..., 2] / 128 - (stats[..., 1] / 128) ** 2
        assert float((mean_g - e.mean(-1)).abs().max()) < 2e-4
        assert float((var_g - e.var(-1, unbiased=False)).abs().max()) < 2e-3


def test_f16c_encoder_bytes_match_the_reference_encoder(cuda_device):
    """The device encoder (split2_f16c) against oracle/f16c_format.py, bit for bit: f16 round-to-nearest, e5m2 residual
    x 2^6 and e5m2 of h x 2^-6, block layout [32 f16 | 32 lo8 | 32 hi8]."""
    from oracle import f16c_format as F16
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(257, 96, generator=g) * torch.logspace(-6, 2, 96)).to(cuda_device)
    got = G.f16c_encode(x)
    exp = F16.encode_rows(x)
    assert got.shape == exp.shape
    bad = (got != exp).nonzero()
    assert bad.numel() == 0, f"{bad.shape[0]} differing bytes, first at {bad[0].tolist()}"
    assert float((F16.decode_rows(got, 96) - x).abs().max() / x.abs().max()) < 2 ** -12


@pytest.mark.parametrize("mode", [4, 2], ids=["bias", "resid"])
@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (1000, 1536, 512), (77, 512, 1024), (4131, 1024, 512)])
def test_linear_f16c_matches_its_own_arithmetic(cuda_device, M, N, K, mode):
    """Tight check of the MMA wiring: the kernel against ah.wh + q(al 2^6) q(wh 2^-6) + q(ah 2^-6) q(wl 2^6) evaluated in
    float64 on the same encoded operands (oracle/f16c_format.py).  Only fp32 accumulation separates the two (~1e-6); a
    mis-paired K-slice or a wrong operand format would show up at 1e-4 ... 1."""
    from oracle import f16c_format as F16
    A, W, b, gamma, beta, resid = _mk(M, N, K, cuda_device, seed=M + N + K + mode)
    y, _ = G.test_linear(mode, A, W, b, gamma, beta, resid, 1e-6, math=2, use_ref=0)
    exp = F16.matmul(A, W) + b.double()
    if mode == 2:
        exp = exp + resid.double()
    rel, mx = _rel(y, exp)
    assert rel < 3e-6, f"rel {rel:.3e} max {mx:.3e}"


@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (1000, 1536, 512), (77, 512, 1024)])
@pytest.mark.parametrize("mode", [4, 0, 2])
def test_linear_bf16_single_pass(cuda_device, M, N, K, mode):
    A, W, b, gamma, beta, resid = _mk(M, N, K, cuda_device, seed=7)
    y, _ = G.test_linear(mode, A, W, b, gamma, beta, resid, 1e-6, math=1, use_ref=0)
    exp = _expected(mode, A, W, b, gamma, beta, resid, 1e-6)
    assert torch.isfinite(y).all()
    rel, mx = _rel(y, exp)
    assert rel < 8e-3, f"rel {rel:.3e} max {mx:.3e}"     # plain bf16 operands: 2^-9 per element


def _attn_expected(qkv, B, F, J, C, H, temporal):
    d = C // H
    q, k, v = qkv.double().reshape(B * F, J, 3, H, d).permute(2, 0, 3, 1, 4)
    if temporal:
        q, k, v = [t.reshape(B, F, H, J, d).permute(0, 2, 3, 1, 4) for t in (q, k, v)]
        att = ((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(-1)
        return (att @ v).permute(0, 3, 2, 1, 4).reshape(B * F * J, C)
    att = ((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(-1)
    return (att @ v).transpose(1, 2).reshape(B * F * J, C)


ATT_SHAPES = [(2, 27, 17, 512, 8), (1, 243, 17, 512, 8), (3, 16, 17, 256, 8), (2, 130, 17, 512, 8), (1, 1, 17, 256, 8),
              (2, 243, 17, 256, 8), (5, 30, 17, 512, 8), (9, 200, 17, 512, 8), (3, 81, 17, 256, 8), (7, 32, 17, 256, 8),
              (1, 5, 17, 512, 8), (6, 16, 17, 512, 8)]


@pytest.mark.parametrize("use_ref", [1, 3, 0], ids=["simt_ref", "tcgen05unpacked", "tcgen05"])
@pytest.mark.parametrize("B,F,J,C,H", ATT_SHAPES)
def test_temporal_attention(cuda_device, B, F, J, C, H, use_ref):
    g = torch.Generator().manual_seed(B * 1000 + F)
    qkv = (torch.randn(B * F * J, 3 * C, generator=g) * 1.2).to(cuda_device)
    y = G.test_attention(1, qkv, B, F, J, C, H, math=0, use_ref=use_ref)
    exp = _attn_expected(qkv, B, F, J, C, H, True)
    assert torch.isfinite(y).all(), "non-finite / unwritten output"
    rel, mx = _rel(y, exp)
    assert rel < 3e-5, f"rel {rel:.3e} max {mx:.3e}"


@pytest.mark.parametrize("use_ref", [1, 0], ids=["simt_ref", "tcgen05"])
@pytest.mark.parametrize("B,F,J,C,H", ATT_SHAPES)
def test_spatial_attention(cuda_device, B, F, J, C, H, use_ref):
    g = torch.Generator().manual_seed(B * 77 + F)
    qkv = (torch.randn(B * F * J, 3 * C, generator=g) * 1.2).to(cuda_device)
    y = G.test_attention(0, qkv, B, F, J, C, H, math=0, use_ref=use_ref)
    exp = _attn_expected(qkv, B, F, J, C, H, False)
    assert torch.isfinite(y).all()
    rel, mx = _rel(y, exp)
    assert rel < 3e-5, f"rel {rel:.3e} max {mx:.3e}"


@pytest.mark.parametrize("use_ref", [3, 0], ids=["f16c_unpacked", "f16c"])
@pytest.mark.parametrize("B,F,J,C,H", ATT_SHAPES)
def test_temporal_attention_f16c(cuda_device, B, F, J, C, H, use_ref):
    """F16C attention (attn_t_f16c.cuh, and attn_s_f16c.cuh in its packed-temporal mode for F <= 32): F16C rows in,
    2 fp16 + 2 e5m2 MMAs per 32-channel block for Q K^T and P V, F16C rows out (decoded as h + l)."""
    g = torch.Generator().manual_seed(B * 1000 + F)
    qkv = (torch.randn(B * F * J, 3 * C, generator=g) * 1.2).to(cuda_device)
    y = G.test_attention(1, qkv, B, F, J, C, H, math=2, use_ref=use_ref)
    exp = _attn_expected(qkv, B, F, J, C, H, True)
    assert torch.isfinite(y).all()
    rel, mx = _rel(y, exp)
    print(f"f16c temporal B={B} F={F} C={C}: rel {rel:.3e} max {mx:.3e}")
    assert rel < 2e-4, f"rel {rel:.3e} max {mx:.3e}"


@pytest.mark.parametrize("B,F,J,C,H", ATT_SHAPES)
def test_spatial_attention_f16c(cuda_device, B, F, J, C, H):
    g = torch.Generator().manual_seed(B * 77 + F)
    qkv = (torch.randn(B * F * J, 3 * C, generator=g) * 1.2).to(cuda_device)
    y = G.test_attention(0, qkv, B, F, J, C, H, math=2, use_ref=0)
    exp = _attn_expected(qkv, B, F, J, C, H, False)
    assert torch.isfinite(y).all()
    rel, mx = _rel(y, exp)
    print(f"f16c spatial B={B} F={F} C={C}: rel {rel:.3e} max {mx:.3e}")
    assert rel < 2e-4, f"rel {rel:.3e} max {mx:.3e}"


@pytest.mark.parametrize("B,F,J,C,H", [(2, 27, 17, 512, 8), (1, 243, 17, 256, 8)])
def test_temporal_attention_bf16_single_pass(cuda_device, B, F, J, C, H):
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B * F * J, 3 * C, generator=g).to(cuda_device)
    y = G.test_attention(1, qkv, B, F, J, C, H, math=1, use_ref=0)
    exp = _attn_expected(qkv, B, F, J, C, H, True)
    assert torch.isfinite(y).all()
    rel, mx = _rel(y, exp)
    assert rel < 1.5e-2, f"rel {rel:.3e} max {mx:.3e}"


@pytest.mark.parametrize("M,N,K", [(64, 128, 256), (1000, 512, 512), (4131, 1536, 512), (20000, 512, 1024), (33, 256, 256)])
@pytest.mark.parametrize("math_mode", [0, 1], ids=["bf16x3", "bf16"])
def test_weight_gradient_kernel(cuda_device, M, N, K, math_mode):
    """Groundwork for the native backward: dW = dY^T X (split-K over the tokens, both operands MN-major)."""
    g = torch.Generator().manual_seed(M + N)
    Gm = torch.randn(M, N, generator=g).to(cuda_device)
    X = (torch.randn(M, K, generator=g) * 1.3 + 0.2).to(cuda_device)
    dW = G.test_wgrad(Gm, X, math=math_mode)
    exp = Gm.double().T @ X.double()
    assert torch.isfinite(dW).all()
    rel, mx = _rel(dW, exp)
    assert rel < (3e-5 if math_mode == 0 else 8e-3), f"rel {rel:.3e} max {mx:.3e}"


@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (4131, 1536, 512), (1000, 1024, 512), (77, 512, 1024), (459, 256, 256)])
@pytest.mark.parametrize("math_mode", [0, 1], ids=["bf16x3", "bf16"])
def test_data_gradient_kernel(cuda_device, M, N, K, math_mode):
    """Groundwork for the native backward: dX = dY W with W read in its forward [N][K] layout (MN-major B operand)."""
    g = torch.Generator().manual_seed(M + K)
    Gm = torch.randn(M, N, generator=g).to(cuda_device)
    W = (torch.randn(N, K, generator=g) / math.sqrt(N)).to(cuda_device)
    dX = G.test_dgrad(Gm, W, math=math_mode)
    exp = Gm.double() @ W.double()
    assert torch.isfinite(dX).all()
    rel, mx = _rel(dX, exp)
    assert rel < (3e-5 if math_mode == 0 else 8e-3), f"rel {rel:.3e} max {mx:.3e}"


def _attn_autograd(qkv, dO, B, F, J, C, H, temporal):
    d = C // H
    x = qkv.double().clone().requires_grad_(True)
    q, k, v = x.reshape(B * F, J, 3, H, d).permute(2, 0, 3, 1, 4)
    if temporal:
        q, k, v = [t.reshape(B, F, H, J, d).permute(0, 2, 3, 1, 4) for t in (q, k, v)]
        o = (((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(-1) @ v).permute(0, 3, 2, 1, 4).reshape(B * F * J, C)
    else:
        o = (((q @ k.transpose(-2, -1)) * d ** -0.5).softmax(-1) @ v).transpose(1, 2).reshape(B * F * J, C)
    o.backward(dO.double())
    return x.grad


@pytest.mark.parametrize("temporal", [1, 0], ids=["temporal", "spatial"])
@pytest.mark.parametrize("B,F,J,C,H", [(2, 27, 17, 512, 8), (1, 243, 17, 512, 8), (2, 130, 17, 256, 8), (3, 16, 17, 512, 8),
                                       (1, 1, 17, 256, 8), (3, 9, 17, 256, 8), (1, 32, 5, 256, 4)])
def test_attention_core_backward(cuda_device, B, F, J, C, H, temporal):
    """Groundwork for the native backward: flash-style tcgen05 attention backward (bf16 single pass) vs autograd."""
    g = torch.Generator().manual_seed(B * 31 + F)
    qkv = torch.randn(B * F * J, 3 * C, generator=g).to(cuda_device)
    dO = torch.randn(B * F * J, C, generator=g).to(cuda_device)
    # the kernels see bf16-rounded inputs: compare against autograd on the same rounded values
    qkv_r, dO_r = qkv.bfloat16().float(), dO.bfloat16().float()
    got = G.test_attention_backward(temporal, qkv, dO, B, F, J, C, H)
    exp = _attn_autograd(qkv_r, dO_r, B, F, J, C, H, bool(temporal))
    assert torch.isfinite(got).all(), "non-finite / unwritten gradient"
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        rel, mx = _rel(got[:, sl], exp[:, sl])
        # F = 1: softmax of a single key is constant -> dq = dk = 0 exactly, only an absolute check makes sense
        assert rel < 2e-2 or mx < 1e-5, f"{name}: rel {rel:.3e} max {mx:.3e}"
