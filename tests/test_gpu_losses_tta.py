"""GPU parity for the fused pretrain-loss kernel (row f1) and the one-call flip TTA (row f2)."""
import os

import numpy as np
import pytest
import torch

from conftest import build_module
from motionbert_b200 import loss as ML
from motionbert_b200 import tta
from oracle import dstformer_oracle as O
from oracle import pretrain_loss_oracle as LO

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pretrain_loss.npz")


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_fused_loss_matches_reference_fixture(cuda_device, case):
    z = np.load(GOLD)
    p = torch.from_numpy(z[f"{case}_pred"]).to(cuda_device).requires_grad_(True)
    g = torch.from_numpy(z[f"{case}_target"]).to(cuda_device)
    total, parts = ML.pretrain_loss_3d(p, g, 0.5, 20.0)
    total.backward()
    exp = z[f"{case}_losses"]
    assert np.allclose(parts.cpu().numpy(), exp, rtol=2e-6, atol=1e-7), (parts.cpu().numpy(), exp)
    assert abs(float(total) - exp[3]) <= 2e-6 * abs(exp[3])
    gr = torch.from_numpy(z[f"{case}_grad"]).to(cuda_device)
    assert float((p.grad.double() - gr).abs().max()) <= 2e-5 * float(gr.abs().max())
    p2 = torch.from_numpy(z[f"{case}_pred"]).to(cuda_device).requires_grad_(True)
    l2d = ML.loss_2d_weighted(p2, g, torch.from_numpy(z[f"{case}_conf"]).to(cuda_device))
    l2d.backward()
    assert abs(float(l2d) - float(z[f"{case}_loss2d"])) <= 2e-6 * float(z[f"{case}_loss2d"])
    g2 = torch.from_numpy(z[f"{case}_grad2d"]).to(cuda_device)
    assert float((p2.grad.double() - g2).abs().max()) <= 2e-5 * float(g2.abs().max())


def test_fused_loss_at_training_size_and_grad_scaling(cuda_device):
    p_np, g_np, _ = LO.make_case(16, 243, 17, 9)
    p = torch.from_numpy(p_np).to(cuda_device).requires_grad_(True)
    g = torch.from_numpy(g_np).to(cuda_device)
    total, parts = ML.pretrain_loss_3d(p, g, 0.5, 20.0)
    (3.0 * total).backward()
    tot_o, parts_o = LO.pretrain_total(p_np, g_np, 0.5, 20.0)
    assert abs(float(total) - tot_o) <= 3e-6 * tot_o
    assert np.allclose(parts.cpu().numpy()[:3], parts_o, rtol=3e-6)
    po = torch.from_numpy(p_np).double().requires_grad_(True)
    to, _ = LO.torch_total(po, torch.from_numpy(g_np).double(), 0.5, 20.0)
    (go,) = torch.autograd.grad(3.0 * to, po)
    assert float((p.grad.cpu().double() - go).abs().max()) <= 3e-5 * float(go.abs().max())
    # individual reference-named entry points
    assert abs(float(ML.loss_mpjpe(p.detach(), g)) - parts_o[0]) <= 3e-6 * parts_o[0]
    assert abs(float(ML.n_mpjpe(p.detach(), g)) - parts_o[1]) <= 3e-6 * parts_o[1]
    assert abs(float(ML.loss_velocity(p.detach(), g)) - parts_o[2]) <= 3e-6 * parts_o[2]
    with pytest.raises(RuntimeError):
        ML.loss_mpjpe(p.detach().cpu(), g.cpu())


def test_flip_tta_single_call_equals_two_calls(cuda_device):
    cfg = O.LITE
    m = build_module(cfg, O.make_params(cfg, 2), cuda_device)
    x = torch.from_numpy(O.make_input(3, 27, 17, 4)).to(cuda_device)
    with torch.no_grad():
        one = tta.forward_flip_tta(m, x)
        two = (m(x) + tta.flip_data(m(tta.flip_data(x)))) * 0.5
    assert one.shape == (3, 27, 17, 3)
    assert float((one - two).abs().max()) <= 1e-6 * float(two.abs().max())


def test_training_step_with_fused_loss(cuda_device):
    """model -> fused loss -> native backward: the whole device-side pretrain step without a host sync."""
    from test_gpu_backward import _module
    m = _module(cuda_device, 256, 1, 8, 2, seed=4)
    x = torch.from_numpy(O.make_input(2, 12, 17, 3)).to(cuda_device)
    gt = torch.from_numpy(O.make_input(2, 12, 17, 5)).to(cuda_device)
    total, parts = ML.pretrain_loss_3d(m(x), gt, 0.5, 20.0)
    total.backward()
    ref_total, _ = LO.torch_total(m(x).detach().cpu().double(), gt.cpu().double(), 0.5, 20.0)
    assert abs(float(total) - float(ref_total)) < 1e-4 * float(ref_total)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert float(m.head.weight.grad.abs().sum()) > 0
