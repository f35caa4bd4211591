"""Differentiable torch restatement of the DSTformer forward  --  TEST INFRASTRUCTURE ONLY.

Same op sequence as `oracle/dstformer_torch_cpu.py` (reference: lib/model/DSTformer.py:329-358, Block :239-249,
Attention :138-200, MLP :79-85, DropPath lib/model/drop.py:17-32 as a per-frame scale vector), but taking the 260
parameters as a flat list in `DSTformer._ordered_params()` order and written for autograd in any dtype/device, so that
float64 autograd through it is the checker of the hand-written backward kernels (`mb_backward`).  It is pinned against
gradients of the REAL reference module by `oracle/make_golden_grads.py` -> tests/golden/grads_*.npz
(tests/test_oracle.py::test_torch_autograd_restatement_matches_reference_gradients).
Only tests/, `__graft_entry__.smoke()` and bench.py's checker legs may import it; the product package never does.
"""
from __future__ import annotations

import torch
import torch.nn.functional as Fn


def _attention(x, qkv_w, qkv_b, proj_w, proj_b, temporal, F, H, scale):
    BF, J, C = x.shape
    d = C // H
    qkv = Fn.linear(x, qkv_w, qkv_b).reshape(BF, J, 3, H, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if temporal:
        q = q.reshape(-1, F, H, J, d).permute(0, 2, 3, 1, 4)
        k = k.reshape(-1, F, H, J, d).permute(0, 2, 3, 1, 4)
        v = v.reshape(-1, F, H, J, d).permute(0, 2, 3, 1, 4)
        o = Fn.scaled_dot_product_attention(q, k, v, scale=scale)
        o = o.permute(0, 3, 2, 1, 4).reshape(BF, J, C)
    else:
        o = Fn.scaled_dot_product_attention(q, k, v, scale=scale)
        o = o.transpose(1, 2).reshape(BF, J, C)
    return Fn.linear(o, proj_w, proj_b)


def recompute_forward(mod, x, return_rep, dp_scale, P):
    """Differentiable torch-op restatement of the forward used only inside backward."""
    B, F, J, _ = x.shape
    C, H = mod.dim_feat, mod.num_heads
    scale = mod.qk_scale or (C // H) ** -0.5
    it = iter(P)
    nxt = lambda: next(it)   # noqa: E731
    temp, pos, je_w, je_b = nxt(), nxt(), nxt(), nxt()
    blocks = []
    for _s in range(2):
        for _i in range(mod.depth):
            blocks.append([nxt() for _ in range(24)])
    norm_w, norm_b, pl_w, pl_b, head_w, head_b = nxt(), nxt(), nxt(), nxt(), nxt(), nxt()
    ts = [(nxt(), nxt()) for _ in range(mod.depth)] if mod.att_fuse else None

    h = Fn.linear(x.reshape(B * F, J, -1), je_w, je_b) + pos
    h = (h.reshape(B, F, J, C) + temp[:, :F]).reshape(B * F, J, C)

    def run_block(h, w, order, sub0):
        (n1s_w, n1s_b, n1t_w, n1t_b, ps_w, ps_b, qs_w, qs_b, pt_w, pt_b, qt_w, qt_b,
         n2s_w, n2s_b, n2t_w, n2t_b, f1s_w, f1s_b, f2s_w, f2s_b, f1t_w, f1t_b, f2t_w, f2t_b) = w
        sub = sub0
        for which in order:
            if which == "S":
                n1, n2 = (n1s_w, n1s_b), (n2s_w, n2s_b)
                att, mlp = (qs_w, qs_b, ps_w, ps_b), (f1s_w, f1s_b, f2s_w, f2s_b)
            else:
                n1, n2 = (n1t_w, n1t_b), (n2t_w, n2t_b)
                att, mlp = (qt_w, qt_b, pt_w, pt_b), (f1t_w, f1t_b, f2t_w, f2t_b)
            y = _attention(Fn.layer_norm(h, (C,), n1[0], n1[1], mod.eps), *att, which == "T", F, H, scale)
            if dp_scale is not None:
                y = y * dp_scale[sub].view(-1, 1, 1)
            h = h + y
            sub += 1
            y = Fn.layer_norm(h, (C,), n2[0], n2[1], mod.eps)
            y = Fn.linear(Fn.gelu(Fn.linear(y, mlp[0], mlp[1])), mlp[2], mlp[3])
            if dp_scale is not None:
                y = y * dp_scale[sub].view(-1, 1, 1)
            h = h + y
            sub += 1
        return h

    for i in range(mod.depth):
        x_st = run_block(h, blocks[i], "ST", i * 8)
        x_ts = run_block(h, blocks[mod.depth + i], "TS", i * 8 + 4)
        if ts is not None:
            a = Fn.linear(torch.cat([x_st, x_ts], dim=-1), ts[i][0], ts[i][1]).softmax(dim=-1)
            h = x_st * a[:, :, 0:1] + x_ts * a[:, :, 1:2]
        else:
            h = (x_st + x_ts) * 0.5
    h = Fn.layer_norm(h, (C,), norm_w, norm_b, mod.eps).reshape(B, F, J, C)
    rep = torch.tanh(Fn.linear(h, pl_w, pl_b))
    return rep if return_rep else Fn.linear(rep, head_w, head_b)
