"""Pretrain-step losses of train.py:178-199 on the GPU, fused with their gradient.

Same names and argument meaning as lib/model/loss.py (`loss_mpjpe`, `n_mpjpe`, `loss_velocity`, `loss_2d_weighted`),
plus `pretrain_loss_3d`, the weighted sum the training loop builds.  One kernel launch reads the (B, T, J, 3) pose
output once and writes the three loss values AND d(total)/d(pred); `backward` only scales that gradient.  The loss
values stay on the device (no `.item()`): read them when you log, not every step.  CUDA tensors only."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _run(pred, target, conf, lambda_scale, lambda_velocity, want_grad):
    if not pred.is_cuda:
        raise RuntimeError("motionbert_b200.loss runs on sm_100a CUDA devices only (no CPU fallback)")
    if pred.shape != target.shape or pred.dim() != 4 or pred.shape[-1] != 3:
        raise ValueError(f"expected matching (B, T, J, 3) tensors, got {tuple(pred.shape)} and {tuple(target.shape)}")
    B, T, J, _ = pred.shape
    p = pred.detach().float().contiguous()
    g = target.detach().float().contiguous()
    c = None
    if conf is not None:
        if conf.numel() != B * T * J:
            raise ValueError(f"conf must hold B*T*J = {B * T * J} confidences, got {tuple(conf.shape)}")
        c = conf.detach().float().contiguous()
    losses = torch.empty(4, dtype=torch.float32, device=p.device)
    d_pred = torch.empty_like(p) if want_grad else None
    scratch = torch.empty(4, dtype=torch.float64, device=p.device)
    lib = _lib.load()
    with torch.cuda.device(p.device):
        _lib.check(lib.mb_pretrain_loss(p.data_ptr(), g.data_ptr(), c.data_ptr() if c is not None else None, B, T, J,
                                        ctypes.c_float(lambda_scale), ctypes.c_float(lambda_velocity), losses.data_ptr(),
                                        d_pred.data_ptr() if d_pred is not None else None, scratch.data_ptr(),
                                        torch.cuda.current_stream(p.device).cuda_stream), "mb_pretrain_loss")
    return losses, d_pred


class _PoseLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, conf, lambda_scale, lambda_velocity):
        losses, d_pred = _run(pred, target, conf, lambda_scale, lambda_velocity, pred.requires_grad)
        ctx.save_for_backward(d_pred) if d_pred is not None else None
        ctx.has_grad = d_pred is not None
        ctx.mark_non_differentiable(losses)
        return losses[3].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g_parts):
        if not ctx.has_grad:
            return None, None, None, None, None
        (d_pred,) = ctx.saved_tensors
        return d_pred * g_total, None, None, None, None


def pretrain_loss_3d(predicted, target, lambda_scale=0.5, lambda_velocity=20.0):
    """total = loss_mpjpe + lambda_scale * n_mpjpe + lambda_velocity * loss_velocity (train.py:178-191 with the shipped
    configs' zero limb / angle weights, MB_pretrain.yaml:39-44).  Returns (total, parts) where total is differentiable
    w.r.t. `predicted` and parts = device tensor [loss_3d_pos, loss_3d_scale, loss_3d_velocity, total]."""
    return _PoseLoss.apply(predicted, target, None, float(lambda_scale), float(lambda_velocity))


def loss_mpjpe(predicted, target):                     # lib/model/loss.py:56-63
    return _PoseLoss.apply(predicted, target, None, 0.0, 0.0)[0]


def n_mpjpe(predicted, target):                        # lib/model/loss.py:80-89 (value only; use pretrain_loss_3d to train)
    return _run(predicted, target, None, 0.0, 0.0, False)[0][1]


def loss_velocity(predicted, target):                  # lib/model/loss.py:133-142 (value only)
    return _run(predicted, target, None, 0.0, 0.0, False)[0][2]


def loss_2d_weighted(predicted, target, conf):         # lib/model/loss.py:73-78
    return _PoseLoss.apply(predicted, target, conf, 0.0, 0.0)[0]
