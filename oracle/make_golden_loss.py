"""Generates tests/golden/pretrain_loss.npz from the REAL reference (lib/model/loss.py imported from /root/reference in
the build container): loss values and autograd gradients of the weighted pretrain objective for seeded inputs, and
checks the oracle restatement against them.  Run here only; the fixture travels, /root/reference does not."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from lib.model import loss as R  # noqa: E402
from oracle import pretrain_loss_oracle as O  # noqa: E402

out = {}
for name, (B, T, J, seed) in {"a": (3, 9, 17, 1), "b": (2, 1, 17, 2), "c": (1, 40, 17, 3)}.items():
    p, g, conf = O.make_case(B, T, J, seed)
    pt = torch.from_numpy(p).double().requires_grad_(True)
    gt = torch.from_numpy(g).double()
    l1, l2, l3 = R.loss_mpjpe(pt, gt), R.n_mpjpe(pt, gt), R.loss_velocity(pt, gt)
    total = l1 + 0.5 * l2 + 20.0 * l3.to(l1.dtype)
    (grad,) = torch.autograd.grad(total, pt)
    pt2 = torch.from_numpy(p).double().requires_grad_(True)
    l2d = R.loss_2d_weighted(pt2, gt, torch.from_numpy(conf).double())
    (grad2d,) = torch.autograd.grad(l2d, pt2)
    # oracle vs reference
    tot_o, parts_o = O.pretrain_total(p, g, 0.5, 20.0)
    assert abs(tot_o - float(total)) < 1e-12 and abs(parts_o[1] - float(l2)) < 1e-12, (tot_o, float(total))
    assert abs(O.loss_2d_weighted(p, g, conf) - float(l2d)) < 1e-12
    po = torch.from_numpy(p).double().requires_grad_(True)
    to, _ = O.torch_total(po, gt, 0.5, 20.0)
    (go,) = torch.autograd.grad(to, po)
    assert float((go - grad).abs().max()) < 1e-12
    out.update({f"{name}_pred": p, f"{name}_target": g, f"{name}_conf": conf,
                f"{name}_losses": np.array([float(l1), float(l2), float(l3), float(total)]),
                f"{name}_grad": grad.numpy(), f"{name}_loss2d": np.array(float(l2d)), f"{name}_grad2d": grad2d.numpy()})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pretrain_loss.npz"), **out)
print("wrote tests/golden/pretrain_loss.npz; oracle == reference to 1e-12")
