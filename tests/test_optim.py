"""Row f4: AdamW over the encoder's parameter table in grouped launches + grouped re-pack."""
from functools import partial

import pytest
import torch
import torch.nn as nn

from oracle import dstformer_oracle as O


def _model(dev=None, depth=2):
    from motionbert_b200 import DSTformer
    torch.manual_seed(0)
    m = DSTformer(dim_in=3, dim_out=3, dim_feat=256, dim_rep=512, depth=depth, num_heads=8, mlp_ratio=2,
                  norm_layer=partial(nn.LayerNorm, eps=1e-6))
    return m.to(dev) if dev is not None else m


def test_state_dict_layout_is_torch_adamw():
    from motionbert_b200.optim import AdamW
    m = _model()
    ours = AdamW(m, lr=5e-4, weight_decay=0.01)
    ref = torch.optim.AdamW(m.parameters(), lr=5e-4, weight_decay=0.01)
    a, b = ours.state_dict(), ref.state_dict()
    assert a["param_groups"][0]["params"] == b["param_groups"][0]["params"]
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert a["param_groups"][0][k] == b["param_groups"][0][k]
    ours.load_state_dict(b)                                   # a torch.optim.AdamW checkpoint loads (train.py:271-276)
    for g in ours.param_groups:
        g["lr"] *= 0.99                                       # train.py:360-363 lr decay


@pytest.mark.gpu
def test_native_adamw_matches_torch_and_repacks(cuda_device):
    from motionbert_b200.optim import AdamW
    m1, m2 = _model(cuda_device).train(), _model(cuda_device).train()
    m2.load_state_dict(m1.state_dict())
    head = nn.Linear(8, 4).to(cuda_device)                    # a tensor outside the encoder (task head)
    head2 = nn.Linear(8, 4).to(cuda_device)
    head2.load_state_dict(head.state_dict())
    for p in m1.blocks_st[0].attn_t.parameters():             # partial_train: frozen tensors are skipped
        p.requires_grad_(False)
    for p in m2.blocks_st[0].attn_t.parameters():
        p.requires_grad_(False)
    o1 = AdamW(m1, [{"params": [p for p in m1.parameters() if p.requires_grad], "lr": 1e-3},
                    {"params": head.parameters(), "lr": 1e-2}], weight_decay=0.05)
    o2 = torch.optim.AdamW([{"params": [p for p in m2.parameters() if p.requires_grad], "lr": 1e-3},
                            {"params": head2.parameters(), "lr": 1e-2}], weight_decay=0.05)
    x = torch.from_numpy(O.make_input(2, 12, 17, 3)).to(cuda_device)
    tgt = torch.from_numpy(O.make_input(2, 12, 17, 4)).to(cuda_device)
    losses = []
    for it in range(4):
        for m, o, h in ((m1, o1, head), (m2, o2, head2)):
            o.zero_grad(set_to_none=True)
            loss = ((m(x) - tgt) ** 2).mean() + h(torch.ones(1, 8, device=cuda_device)).pow(2).sum()
            loss.backward()
            o.step()
            losses.append(float(loss.detach()))
    # the two models see slightly different gradients from step 2 on only through rounding of the updates themselves
    for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(p1, p2, rtol=2e-4, atol=2e-6), n
    for p1, p2 in zip(head.parameters(), head2.parameters()):
        assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-7)
    assert abs(losses[-2] - losses[-1]) < 1e-3 * abs(losses[-1])      # the forward picked the updated weights up (re-pack)
    assert losses[-2] < losses[0]
    sd = o1.state_dict()
    assert int(sd["state"][0]["step"]) == 4 and sd["state"][0]["exp_avg"].shape == m1.temp_embed.shape
    o2.load_state_dict(sd)                                    # and the state goes back into a plain torch.optim.AdamW
