"""Reference encoder / decoder / product of the "F16C" operand format  --  TEST INFRASTRUCTURE ONLY.

The format is this repo's own (motionbert_b200/csrc/ptx.cuh; it has no counterpart in the reference, whose arithmetic is
plain fp32): x ~= h + l with h = f16_rn(x); per 32 consecutive elements one 128-byte block
    [ 32 x f16 h | 32 x e5m2 rn(l * 2^6) | 32 x e5m2 rn(h * 2^-6) ]
and a product a.w is evaluated by the tensor cores as  ah.wh + q(al 2^6) q(wh 2^-6) + q(ah 2^-6) q(wl 2^6)  with fp32
accumulation.  These torch restatements (any device) let the tests check the kernels' bytes bit for bit and their GEMM
results against the scheme's own arithmetic (tight), next to the checks against exact float64 (loose, = the scheme's error).
"""
from __future__ import annotations

import torch

S = 64.0


def split(x: torch.Tensor):
    """x (float) -> (h f16, lo8 e5m2 of l*64, hi8 e5m2 of h/64)"""
    x = x.float()
    h = x.half()
    l = x - h.float()
    lo8 = (l * S).to(torch.float8_e5m2)
    hi8 = (h * torch.tensor(1.0 / S, dtype=torch.float16, device=x.device)).to(torch.float8_e5m2)
    return h, lo8, hi8


def encode_rows(x: torch.Tensor) -> torch.Tensor:
    """[rows, cols] (cols % 32 == 0) -> uint8 [rows, cols * 4] in the kernels' block layout."""
    rows, cols = x.shape
    h, lo8, hi8 = split(x)
    hb = h.view(torch.uint8).reshape(rows, cols // 32, 64)
    lb = lo8.view(torch.uint8).reshape(rows, cols // 32, 32)
    gb = hi8.view(torch.uint8).reshape(rows, cols // 32, 32)
    return torch.cat([hb, lb, gb], dim=-1).reshape(rows, cols * 4).contiguous()


def decode_rows(b: torch.Tensor, cols: int) -> torch.Tensor:
    """inverse of encode_rows up to the format's precision: h + l"""
    rows = b.shape[0]
    blk = b.reshape(rows, cols // 32, 128)
    h = blk[..., :64].contiguous().view(torch.float16).reshape(rows, cols).float()
    l = blk[..., 64:96].contiguous().view(torch.float8_e5m2).reshape(rows, cols).float() / S
    return h + l


def matmul(a: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """a [M,K] . w [N,K]^T exactly as the F16C MMAs see the operands (float64 accumulation)."""
    ah, al8, ah8 = split(a)
    wh, wl8, wh8 = split(w)
    d = torch.float64
    return ah.to(d) @ wh.to(d).T + al8.to(d) @ wh8.to(d).T + ah8.to(d) @ wl8.to(d).T
