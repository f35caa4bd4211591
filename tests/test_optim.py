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
    """mb_adamw_step against torch.optim.AdamW on IDENTICAL gradients (the native backward's atomics make two separately
    computed gradients differ in the last bits, which Adam's g / sqrt(v) amplifies), four steps, two param groups, frozen
    tensors, a head outside the encoder; then the forward must see the updated weights (grouped re-pack)."""
    from motionbert_b200.optim import AdamW
    m1, m2 = _model(cuda_device).train(), _model(cuda_device).train()
    m2.load_state_dict(m1.state_dict())
    head = nn.Linear(8, 4).to(cuda_device)                    # a tensor outside the encoder (task head)
    head2 = nn.Linear(8, 4).to(cuda_device)
    head2.load_state_dict(head.state_dict())
    for m in (m1, m2):
        for p in m.blocks_st[0].attn_t.parameters():          # partial_train: frozen tensors are skipped
            p.requires_grad_(False)
    o1 = AdamW(m1, [{"params": [p for p in m1.parameters() if p.requires_grad], "lr": 1e-3},
                    {"params": head.parameters(), "lr": 1e-2}], weight_decay=0.05)
    o2 = torch.optim.AdamW([{"params": [p for p in m2.parameters() if p.requires_grad], "lr": 1e-3},
                            {"params": head2.parameters(), "lr": 1e-2}], weight_decay=0.05)
    x = torch.from_numpy(O.make_input(2, 12, 17, 3)).to(cuda_device)
    tgt = torch.from_numpy(O.make_input(2, 12, 17, 4)).to(cuda_device)
    losses = []
    for it in range(4):
        o1.zero_grad(set_to_none=True)
        loss = ((m1(x) - tgt) ** 2).mean() + head(torch.ones(1, 8, device=cuda_device)).pow(2).sum()
        loss.backward()
        losses.append(float(loss.detach()))
        for p1, p2 in zip(list(m1.parameters()) + list(head.parameters()), list(m2.parameters()) + list(head2.parameters())):
            p2.grad = None if p1.grad is None else p1.grad.clone()
        o1.step()
        o2.step()
        for (n, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
            assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-7), (it, n, float((p1 - p2).abs().max()))
    for p1, p2 in zip(head.parameters(), head2.parameters()):
        assert torch.equal(p1, p2)
    assert losses[-1] < losses[0]
    # the kernels wrote the weights behind autograd's back: the next forward must run on re-packed operands
    m3 = _model(cuda_device).eval()
    m3.load_state_dict(m1.state_dict())
    with torch.no_grad():
        assert torch.equal(m1.eval()(x), m3(x))
    sd = o1.state_dict()
    assert int(sd["state"][0]["step"]) == 4 and sd["state"][0]["exp_avg"].shape == m1.temp_embed.shape
    o2.load_state_dict(sd)                                    # and the state goes back into a plain torch.optim.AdamW
