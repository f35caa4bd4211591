"""torch-CPU restatement of the DSTformer forward  --  TEST INFRASTRUCTURE ONLY.

Same algorithm as `oracle/dstformer_oracle.py`, written with torch CPU ops
(Linear / layer_norm / gelu / softmax / matmul) in the same op sequence as the
reference module (`/root/reference/lib/model/DSTformer.py:329-358`), so that
timing it on the GPU box's host cores reproduces what the reference's own CPU
forward costs (the reference is PyTorch-only Python and cannot travel to the
GPU box).  Used by `bench.py` for the `cpu_baseline` leg and `--impl reference`,
and by tests as a second checker.  Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as Fn


def _attention(x, P, p, mode, F, H):
    BF, J, C = x.shape
    d = C // H
    scale = d ** -0.5                                                          # DSTformer.py:94
    qkv = Fn.linear(x, P[p + ".qkv.weight"], P[p + ".qkv.bias"])
    qkv = qkv.reshape(BF, J, 3, H, d).permute(2, 0, 3, 1, 4)                   # :143
    q, k, v = qkv[0], qkv[1], qkv[2]
    if mode == "spatial":                                                      # :178-186
        att = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
        o = (att @ v).transpose(1, 2).reshape(BF, J, C)
    else:                                                                      # :188-200
        qt = q.reshape(-1, F, H, J, d).permute(0, 2, 3, 1, 4)
        kt = k.reshape(-1, F, H, J, d).permute(0, 2, 3, 1, 4)
        vt = v.reshape(-1, F, H, J, d).permute(0, 2, 3, 1, 4)
        att = ((qt @ kt.transpose(-2, -1)) * scale).softmax(dim=-1)
        o = (att @ vt).permute(0, 3, 2, 1, 4).reshape(BF, J, C)
    return Fn.linear(o, P[p + ".proj.weight"], P[p + ".proj.bias"])            # :148


def _block(x, P, p, order, F, H, eps):
    C = x.shape[-1]
    for which in order:                                                        # :240-249
        s = "s" if which == "S" else "t"
        h = Fn.layer_norm(x, (C,), P[p + f"norm1_{s}.weight"], P[p + f"norm1_{s}.bias"], eps)
        x = x + _attention(h, P, p + f"attn_{s}", "spatial" if which == "S" else "temporal", F, H)
        h = Fn.layer_norm(x, (C,), P[p + f"norm2_{s}.weight"], P[p + f"norm2_{s}.bias"], eps)
        h = Fn.gelu(Fn.linear(h, P[p + f"mlp_{s}.fc1.weight"], P[p + f"mlp_{s}.fc1.bias"]))   # :79-81
        x = x + Fn.linear(h, P[p + f"mlp_{s}.fc2.weight"], P[p + f"mlp_{s}.fc2.bias"])       # :83
    return x


@torch.no_grad()
def forward(P: "dict[str, torch.Tensor]", x: torch.Tensor, depth: int, num_heads: int,
            eps: float = 1e-6):
    """Returns (out, rep) like `oracle.dstformer_oracle.forward`."""
    B, F, J, _ = x.shape
    C = P["joints_embed.weight"].shape[0]
    h = Fn.linear(x.reshape(B * F, J, -1), P["joints_embed.weight"], P["joints_embed.bias"])  # :333
    h = h + P["pos_embed"]                                                                     # :334
    h = (h.reshape(B, F, J, C) + P["temp_embed"][:, :F]).reshape(B * F, J, C)                 # :336-337
    for i in range(depth):                                                                     # :340-349
        x_st = _block(h, P, f"blocks_st.{i}.", "ST", F, num_heads, eps)
        x_ts = _block(h, P, f"blocks_ts.{i}.", "TS", F, num_heads, eps)
        a = Fn.linear(torch.cat([x_st, x_ts], dim=-1), P[f"ts_attn.{i}.weight"], P[f"ts_attn.{i}.bias"])
        a = a.softmax(dim=-1)
        h = x_st * a[:, :, 0:1] + x_ts * a[:, :, 1:2]
    h = Fn.layer_norm(h, (C,), P["norm.weight"], P["norm.bias"], eps).reshape(B, F, J, C)      # :352-353
    rep = torch.tanh(Fn.linear(h, P["pre_logits.fc.weight"], P["pre_logits.fc.bias"]))         # :354
    out = Fn.linear(rep, P["head.weight"], P["head.bias"])                                     # :357
    return out, rep
