"""Row f3: the action heads consuming the representation through its temporal mean
(lib/model/model_action.py:15-29, :31-48, :62-71), with the mean fused into the encoder's tail epilogue."""
import numpy as np
import pytest
import torch

from conftest import build_module
from oracle import dstformer_oracle as O


def test_action_net_state_dict_matches_the_reference_layout():
    """Same sub-module names / shapes / order as lib/model/model_action.py, so reference checkpoints load strictly."""
    from motionbert_b200.action import ActionNet
    bb = build_module(O.LITE)
    ours = ActionNet(bb, dim_rep=512, num_classes=60, dropout_ratio=0.5, version='class', hidden_dim=2048, num_joints=17)
    keys = [k for k in ours.state_dict().keys() if k.startswith("head.")]
    assert keys == ["head.bn.weight", "head.bn.bias", "head.bn.running_mean", "head.bn.running_var",
                    "head.bn.num_batches_tracked", "head.fc1.weight", "head.fc1.bias", "head.fc2.weight", "head.fc2.bias"]
    assert ours.head.fc1.weight.shape == (2048, 512 * 17) and ours.head.fc2.weight.shape == (60, 2048)
    emb = ActionNet(bb, version='embed', hidden_dim=2048)
    assert [k for k in emb.state_dict().keys() if k.startswith("head.")] == ["head.fc1.weight", "head.fc1.bias"]
    with pytest.raises(Exception, match="Version"):
        ActionNet(bb, version='nope')


def test_heads_match_reference_math_on_cpu():
    """forward(feat) is the reference op sequence (permute, mean over T, mean over M, fc / bn / relu / fc)."""
    from motionbert_b200.action import ActionHeadClassification, ActionHeadEmbed
    torch.manual_seed(0)
    feat = torch.randn(3, 2, 5, 17, 64)
    h = ActionHeadClassification(dim_rep=64, num_classes=7, hidden_dim=32).eval()
    f = feat.permute(0, 1, 3, 4, 2).mean(dim=-1).reshape(3, 2, -1).mean(dim=1)
    exp = h.fc2(torch.relu(h.bn(h.fc1(f))))
    torch.testing.assert_close(h(feat), exp)
    e = ActionHeadEmbed(dim_rep=64, hidden_dim=32).eval()
    torch.testing.assert_close(e(feat), torch.nn.functional.normalize(e.fc1(f), dim=-1))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["f16c", "bf16x3"])
@pytest.mark.parametrize("cfg_name,N,M,T", [("lite", 2, 2, 27), ("base", 1, 2, 243), ("lite", 3, 1, 40)])
def test_pooled_tail_equals_mean_of_representation(cuda_device, cfg_name, N, M, T, mode):
    """mb_forward_pooled == get_representation(x).mean(dim=1), and ActionNet through the pooled tail == ActionNet through
    the full representation (the reference's own op sequence), for the train_action.py input layout (N, M, T, 17, 3)."""
    from motionbert_b200.action import ActionNet
    cfg = O.BASE if cfg_name == "base" else O.LITE
    bb = build_module(cfg, O.make_params(cfg, 5), cuda_device).set_math_mode(mode)
    x = torch.from_numpy(O.make_input(N * M, T, 17, 3)).to(cuda_device)
    with torch.no_grad():
        rep = bb.get_representation(x)
        pooled = bb.get_representation_pooled(x)
    assert pooled.shape == (N * M, 17, 512)
    ref = rep.double().mean(dim=1)
    err = float((pooled.double() - ref).abs().max())
    assert err < 2e-6, err                                    # same products, fp32 accumulation order differs
    torch.manual_seed(1)
    net = ActionNet(bb, version='class', num_classes=60).to(cuda_device).eval()
    x5 = x.reshape(N, M, T, 17, 3)
    with torch.no_grad():
        fused = net(x5)
        feat = bb.get_representation(x).reshape(N, M, T, 17, -1)
        plain = net.head(feat)
    assert fused.shape == (N, 60)
    assert float((fused - plain).abs().max()) < 1e-4 * max(1.0, float(plain.abs().max()))
    # training the backbone: the plain path runs (a gradient has to flow), and it does
    if N > 1:                                                 # BatchNorm1d needs more than one sample in training mode
        net.train()
        out = net(x5)
        out.sum().backward()
        assert bb.blocks_st[0].mlp_s.fc1.weight.grad is not None
