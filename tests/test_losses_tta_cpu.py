"""CPU checks for the rows that sit either side of the path: the loss oracle against the
fixture generated from the real reference, and flip_data against its definition."""
import os

import numpy as np
import pytest
import torch

from motionbert_b200 import tta
from oracle import pretrain_loss_oracle as LO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pretrain_loss.npz")


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_loss_oracle_matches_reference_fixture(case):
    z = np.load(GOLD)
    p, g, conf = z[f"{case}_pred"], z[f"{case}_target"], z[f"{case}_conf"]
    total, parts = LO.pretrain_total(p, g, 0.5, 20.0)
    exp = z[f"{case}_losses"]
    assert np.allclose([*parts, total], exp, rtol=0, atol=1e-12)
    assert abs(LO.loss_2d_weighted(p, g, conf) - float(z[f"{case}_loss2d"])) < 1e-12
    pt = torch.from_numpy(p).double().requires_grad_(True)
    tot, _ = LO.torch_total(pt, torch.from_numpy(g).double(), 0.5, 20.0)
    (gr,) = torch.autograd.grad(tot, pt)
    assert float((gr - torch.from_numpy(z[f"{case}_grad"])).abs().max()) < 1e-12
    pt2 = torch.from_numpy(p).double().requires_grad_(True)
    (g2,) = torch.autograd.grad(LO.torch_2d(pt2, torch.from_numpy(g).double(), torch.from_numpy(conf).double()), pt2)
    assert float((g2 - torch.from_numpy(z[f"{case}_grad2d"])).abs().max()) < 1e-12


def test_make_case_is_seeded():
    a, b = LO.make_case(2, 3, 17, 5), LO.make_case(2, 3, 17, 5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_flip_data_semantics():
    x = torch.arange(2 * 3 * 17 * 3, dtype=torch.float32).reshape(2, 3, 17, 3)
    f = tta.flip_data(x)
    assert torch.equal(tta.flip_data(f), x)                                  # involution
    assert torch.equal(f[..., 0, :], x[..., 0, :] * torch.tensor([-1.0, 1.0, 1.0]))   # root joint: only x negated
    for l, r in zip(tta.LEFT_JOINTS, tta.RIGHT_JOINTS):
        assert torch.equal(f[..., l, 1:], x[..., r, 1:]) and torch.equal(f[..., l, 0], -x[..., r, 0])
    assert torch.equal(tta.flip_data(x[0]), f[0])                            # [F, 17, D] form
    assert x[0, 0, 4, 0] == 12.0                                             # input untouched
    with pytest.raises(ValueError):
        tta.flip_data(torch.zeros(2, 3, 16, 3))


def test_flip_tta_is_one_call_of_2b_sequences():
    calls = []

    def model(x):
        calls.append(tuple(x.shape))
        return x * 2.0 + 1.0           # flip-equivariant up to the +1 on x (which flips sign)

    x = torch.randn(3, 4, 17, 3)
    y = tta.forward_flip_tta(model, x)
    assert calls == [(6, 4, 17, 3)]
    exp = (model(x) + tta.flip_data(model(tta.flip_data(x)))) * 0.5
    assert torch.allclose(y, exp)


def test_fused_loss_has_no_cpu_fallback_and_checks_shapes():
    from motionbert_b200 import loss as ML
    p, g, conf = (torch.from_numpy(a) for a in LO.make_case(2, 3, 17, 1))
    with pytest.raises(RuntimeError, match="CUDA"):
        ML.pretrain_loss_3d(p, g)
    with pytest.raises(RuntimeError, match="CUDA"):
        ML.loss_2d_weighted(p, g, conf)


def test_forward_flip_tta_rejects_representation_output():
    with pytest.raises(ValueError):
        tta.forward_flip_tta(lambda x: x, torch.zeros(1, 2, 17, 3), return_rep=True)
