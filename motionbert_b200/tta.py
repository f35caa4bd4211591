"""Flip test-time augmentation as ONE forward call.

The reference evaluates `model(x)` and `model(flip_data(x))` as two forwards and averages after flipping back
(train.py:67-72, infer_wild.py:73-78).  Sequences are independent, so the two passes are one 2B-batch forward of the
same kernels: half the launches, twice the rows per launch (B=1 clips stop being launch-bound).  `flip_data` keeps the
reference's name and meaning (lib/utils/utils_data.py:54-66): negate x, swap the left/right joints of the H36M
17-joint skeleton."""
from __future__ import annotations

import torch

LEFT_JOINTS = (4, 5, 6, 11, 12, 13)
RIGHT_JOINTS = (1, 2, 3, 14, 15, 16)


def _flip_perm(num_joints: int, device) -> torch.Tensor:
    perm = list(range(num_joints))
    for l, r in zip(LEFT_JOINTS, RIGHT_JOINTS):
        perm[l], perm[r] = r, l
    return torch.tensor(perm, dtype=torch.long, device=device)


def flip_data(data: torch.Tensor) -> torch.Tensor:
    """Horizontal flip of [N, F, 17, D] or [F, 17, D] poses; returns a new tensor (the input is left untouched)."""
    if data.shape[-2] != 17:
        raise ValueError(f"flip_data expects the 17-joint H36M skeleton, got {data.shape[-2]} joints")
    out = data.index_select(-2, _flip_perm(17, data.device))
    sign = torch.ones(data.shape[-1], dtype=data.dtype, device=data.device)
    sign[0] = -1
    return out * sign


def forward_flip_tta(model, x: torch.Tensor, return_rep: bool = False) -> torch.Tensor:
    """(model(x) + flip_data(model(flip_data(x)))) / 2 with a single 2B-batch forward."""
    if return_rep:
        raise ValueError("flip-TTA averages 3-D poses; the representation is not flip-equivariant")
    B = x.shape[0]
    y = model(torch.cat([x, flip_data(x)], dim=0))
    return (y[:B] + flip_data(y[B:])) * 0.5
