"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the pretrain-step losses.

Restates lib/model/loss.py:56-63 (loss_mpjpe), :73-78 (loss_2d_weighted), :80-89 (n_mpjpe), :133-142 (loss_velocity)
and the weighted sum of train.py:178-191, in numpy (fp64 values) and in differentiable torch-CPU ops (gradients).
Pinned against the real reference functions by oracle/make_golden_loss.py -> tests/golden/pretrain_loss.npz.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
import numpy as np


def loss_mpjpe(p, g):
    return float(np.linalg.norm(p.astype(np.float64) - g.astype(np.float64), axis=-1).mean())


def n_mpjpe(p, g):
    p = p.astype(np.float64)
    g = g.astype(np.float64)
    norm_p = (p * p).sum(axis=3, keepdims=True).mean(axis=2, keepdims=True)
    norm_g = (g * p).sum(axis=3, keepdims=True).mean(axis=2, keepdims=True)
    return loss_mpjpe(norm_g / norm_p * p, g)


def loss_velocity(p, g):
    if p.shape[1] <= 1:
        return 0.0
    p = p.astype(np.float64)
    g = g.astype(np.float64)
    return float(np.linalg.norm((p[:, 1:] - p[:, :-1]) - (g[:, 1:] - g[:, :-1]), axis=-1).mean())


def loss_2d_weighted(p, g, conf):
    d = (p[..., :2].astype(np.float64) - g[..., :2].astype(np.float64)) * conf.astype(np.float64)
    return float(np.linalg.norm(d, axis=-1).mean())


def pretrain_total(p, g, lambda_scale, lambda_velocity):
    parts = (loss_mpjpe(p, g), n_mpjpe(p, g), loss_velocity(p, g))
    return parts[0] + lambda_scale * parts[1] + lambda_velocity * parts[2], parts


# ---- differentiable torch-CPU restatement (gradient oracle)
def torch_total(p, g, lambda_scale, lambda_velocity):
    import torch
    l1 = torch.linalg.vector_norm(p - g, dim=-1).mean()
    s = (g * p).sum(dim=(2, 3), keepdim=True) / (p * p).sum(dim=(2, 3), keepdim=True)
    l2 = torch.linalg.vector_norm(s * p - g, dim=-1).mean()
    if p.shape[1] > 1:
        l3 = torch.linalg.vector_norm((p[:, 1:] - p[:, :-1]) - (g[:, 1:] - g[:, :-1]), dim=-1).mean()
    else:
        l3 = torch.zeros((), dtype=p.dtype)
    return l1 + lambda_scale * l2 + lambda_velocity * l3, (l1, l2, l3)


def torch_2d(p, g, conf):
    import torch
    return torch.linalg.vector_norm((p[..., :2] - g[..., :2]) * conf, dim=-1).mean()


def make_case(B, T, J, seed):
    r = np.random.default_rng(seed)
    g = r.uniform(-1, 1, size=(B, T, J, 3)).astype(np.float32)
    p = (g + 0.3 * r.standard_normal((B, T, J, 3))).astype(np.float32)
    conf = r.uniform(0, 1, size=(B, T, J, 1)).astype(np.float32)
    return p, g, conf
