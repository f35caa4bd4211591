"""autograd glue for the drop-in DSTformer.

Forward = the sm_100a CUDA library (`mb_forward`, or `mb_forward_train` when a gradient is needed).
Backward = `mb_backward`: hand-written tcgen05 data-/weight-gradient GEMMs,
attention-core backward and CUDA-core LayerNorm / GELU / fusion / embed kernels, bf16 single-pass arithmetic.
There is no PyTorch-op fallback: configurations the native backward does not cover raise.
"""
from __future__ import annotations

import torch


class DSTformerFunction(torch.autograd.Function):
    """forward = mb_forward_train, backward = mb_backward (hand-written sm_100a kernels).  DropPath (per-frame scale
    vector), the gradient w.r.t. the pose input and `att_fuse=False` (constant 0.5 / 0.5 fusion) are handled natively."""

    @staticmethod
    def forward(ctx, mod, x, return_rep, dp_scale, *params):
        ctx.mod = mod
        ctx.return_rep = return_rep
        ctx.dp_scale = dp_scale
        mod._check_native_backward(x)
        with torch.no_grad():
            out, rep, saved = mod._launch_train(x, not return_rep, dp_scale)
            # the activation region goes through save_for_backward so that autograd releases it right after the
            # backward (unless retain_graph): two steps' regions never coexist in a train.py-style loop
            ctx.save_for_backward(x, rep, saved)
            ctx.param_versions = tuple(p._version for p in params)
        return rep if return_rep else out

    @staticmethod
    def backward(ctx, grad):
        mod = ctx.mod
        live = [p for p in mod._ordered_params() if p is not None]
        if tuple(p._version for p in live) != ctx.param_versions:
            raise RuntimeError("a DSTformer parameter was modified in place between forward and backward: the saved "
                               "activations belong to the old weights (PyTorch would raise 'modified by an inplace "
                               "operation' here)")
        x, rep, saved = ctx.saved_tensors
        g = grad.contiguous().float()
        if g.data_ptr() & 15:                       # a contiguous slice of a gathered gradient (nn.DataParallel): 16-byte reads
            g = g.clone()
        grads, d_x = mod._launch_backward(x, rep, saved, None if ctx.return_rep else g,
                                          g if ctx.return_rep else None, ctx.dp_scale, ctx.needs_input_grad[1])
        skip = mod._head_param_slots() if ctx.return_rep else ()      # the head is not part of get_representation()
        gp = [gr if (ctx.needs_input_grad[4 + i] and i not in skip) else None for i, gr in enumerate(grads)]
        return (None, d_x, None, None, *gp)
