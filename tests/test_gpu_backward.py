"""Gradient parity of the native backward (mb_forward_train + mb_backward).

Two checkers: (1) gradient fixtures produced by float64 autograd through the REAL reference module
(`oracle/make_golden_grads.py` -> tests/golden/grads_*.npz: per-tensor norm, sum and 512 sampled entries of all 260
gradients + the whole input gradient), and (2) fp64 back-propagation through the oracle's differentiable restatement
(`oracle/dstformer_torch_autograd.py`, itself pinned against those fixtures to ~1e-12 by tests/test_oracle.py) for
shapes without a fixture.  The native backward computes in bf16 single-pass arithmetic with fp32 accumulation -- the
arithmetic of the reference's own mixed-precision training -- so the bar is a per-parameter relative L2 error of a few
percent, with the exact numbers printed per parameter class."""
import os
from functools import partial

import numpy as np
import pytest
import torch
import torch.nn as nn

from motionbert_b200 import DSTformer
from oracle import dstformer_oracle as O
from oracle.dstformer_torch_autograd import recompute_forward

pytestmark = pytest.mark.gpu

REL_L2 = 4e-2        # bf16 single-pass backward vs fp64: per-parameter ||g - g_ref|| / ||g_ref||
COS_MIN = 0.999
JOINT_GATE_REL = 1e-1  # fusion gate (ts_attn.*): cancellation-dominated sums over all tokens, see _compare


def _module(dev, dim_feat, depth, heads, mlp_ratio, dim_rep=512, maxlen=243, seed=0):
    torch.manual_seed(seed)
    m = DSTformer(dim_in=3, dim_out=3, dim_feat=dim_feat, dim_rep=dim_rep, depth=depth, num_heads=heads,
                  mlp_ratio=mlp_ratio, norm_layer=partial(nn.LayerNorm, eps=1e-6), maxlen=maxlen)
    # the reference init leaves biases 0 and LayerNorm at (1, 0): perturb so every gradient path is exercised
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith(".bias") or "norm" in n:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
            if n in ("temp_embed", "pos_embed"):
                p.add_(0.02 * torch.randn(p.shape, generator=g))
    return m.to(dev).train()


def _reference_grads(m, x, w_out, return_rep, dp=None):
    ps = [p.detach().double().requires_grad_(True) for p in m._ordered_params()]
    y = recompute_forward(m, x.double(), return_rep, dp.double() if dp is not None else None, ps)
    loss = (y * w_out.double()).sum()
    grads = torch.autograd.grad(loss, ps, allow_unused=True)       # head.* is unused on the representation path
    return [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, ps)], y.detach()


def _compare(m, grads_ref, label, no_grad=()):
    names = [n for n, _ in m.named_parameters()]
    order = {id(p): i for i, p in enumerate(m._ordered_params())}
    rows = []
    pairs = {}
    for n, p in m.named_parameters():
        gr = grads_ref[order[id(p)]]
        g = p.grad
        if n in no_grad:
            assert g is None, f"{label}: {n} must not receive a gradient"
            continue
        assert g is not None, f"{label}: {n} has no gradient"
        g = g.double()
        assert torch.isfinite(g).all(), f"{label}: non-finite gradient in {n}"
        if n.startswith("ts_attn."):
            # The fusion gate's gradient is a sum over ALL tokens of zero-mean per-token terms (a 2-way softmax:
            # d logit_0 = -d logit_1), i.e. cancellation-dominated; its bias is the weight column of a constant
            # feature.  Judge weight and bias as one vector, with a wider bar (JOINT_GATE_REL).
            pairs.setdefault(n.rsplit(".", 1)[0], []).append((g.flatten(), gr.flatten()))
            continue
        den = float(gr.norm())
        rel = float((g - gr).norm()) / (den + 1e-30)
        cos = float((g * gr).sum()) / (float(g.norm()) * den + 1e-30)
        rows.append((n, rel, cos, den))
    gate_rows = []
    for n, lst in pairs.items():
        g = torch.cat([a for a, _ in lst])
        gr = torch.cat([b for _, b in lst])
        den = float(gr.norm())
        gate_rows.append((n + ".{weight,bias}", float((g - gr).norm()) / (den + 1e-30),
                          float((g * gr).sum()) / (float(g.norm()) * den + 1e-30), den))
    # report: worst 12 and per-class medians
    rows_sorted = sorted(rows + gate_rows, key=lambda r: -r[1])
    print(f"\n[{label}] {len(names)} parameters; worst relative L2 errors:")
    for n, rel, cos, den in rows_sorted[:12]:
        print(f"    {n:44s} rel {rel:.3e}  cos {cos:.6f}  |g_ref| {den:.3e}")
    classes = {}
    for n, rel, cos, den in rows:
        key = ".".join(t for t in n.split(".") if not t.isdigit())
        classes.setdefault(key, []).append(rel)
    print("    per-class median rel:", ", ".join(f"{k}={np.median(v):.1e}" for k, v in sorted(classes.items())))
    bad = [(n, rel, cos) for n, rel, cos, den in rows if den > 0 and (rel > REL_L2 or cos < COS_MIN)]
    bad += [(n, rel, cos) for n, rel, cos, den in gate_rows if den > 0 and (rel > JOINT_GATE_REL or cos < 0.995)]
    assert not bad, f"{label}: {len(bad)} parameters out of tolerance, first: {bad[:5]}"


CASES = [
    # label,        dim_feat, depth, heads, mlp, B, F
    ("lite_d2_f9", 256, 2, 8, 2, 2, 9),          # head_dim 32, packed temporal attention (F <= 32)
    ("lite_d2_f40", 256, 2, 8, 4, 3, 40),        # head_dim 32, one-sequence temporal tiles
    ("base_d1_f243", 512, 1, 8, 4, 1, 243),      # head_dim 64, two query tiles per sequence (BASELINE length)
    ("base_full_f81", 512, 5, 8, 2, 2, 81),      # the shipped DSTformer-base (depth 5, mlp_ratio 2): 260 tensors
    ("lite_full_f27", 256, 5, 8, 4, 2, 27),      # the shipped DSTformer-Lite (depth 5, mlp_ratio 4), packed temporal
]


@pytest.mark.parametrize("label,dim_feat,depth,heads,mlp,B,F", CASES, ids=[c[0] for c in CASES])
def test_native_backward_matches_fp64_autograd(cuda_device, label, dim_feat, depth, heads, mlp, B, F):
    m = _module(cuda_device, dim_feat, depth, heads, mlp, seed=11)
    x = torch.from_numpy(O.make_input(B, F, 17, 21)).to(cuda_device)
    g = torch.Generator().manual_seed(5)
    w_out = torch.randn(B, F, 17, 3, generator=g).to(cuda_device)
    out = m(x)
    (out * w_out).sum().backward()
    torch.cuda.synchronize(cuda_device)
    grads_ref, y_ref = _reference_grads(m, x, w_out, False)
    assert float((out.detach().double() - y_ref).abs().max()) < 1e-3 * float(y_ref.abs().max())
    _compare(m, grads_ref, label)


def test_native_backward_through_get_representation(cuda_device):
    """train_action.py / train_mesh.py back-propagate through get_representation (DSTformer.py:360-361)."""
    m = _module(cuda_device, 256, 2, 8, 2, seed=3)
    B, F = 2, 12
    x = torch.from_numpy(O.make_input(B, F, 17, 8)).to(cuda_device)
    w = torch.randn(B, F, 17, 512, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    rep = m.get_representation(x)
    (rep * w).sum().backward()
    grads_ref, _ = _reference_grads(m, x, w, True)
    # head.* is not part of get_representation(): like the reference's autograd, it gets NO gradient (None, not zeros)
    _compare(m, grads_ref, "rep_path", no_grad=("head.weight", "head.bias"))


def test_gradients_accumulate_and_inplace_weight_update_between_forward_and_backward_raises(cuda_device):
    """.grad accumulates over two backward calls; modifying a weight in place between forward and backward raises
    (the saved activations belong to the old weights -- PyTorch raises in the same situation)."""
    m = _module(cuda_device, 256, 1, 8, 2, seed=9)
    x = torch.from_numpy(O.make_input(2, 10, 17, 4)).to(cuda_device)
    (m(x) ** 2).sum().backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    (m(x) ** 2).sum().backward()
    for n, p in m.named_parameters():
        assert torch.allclose(p.grad, 2 * g1[n], rtol=1e-3, atol=1e-5 * float(g1[n].abs().max()) + 1e-12), n
    m.zero_grad(set_to_none=True)
    loss = (m(x) ** 2).sum()
    with torch.no_grad():
        m.blocks_st[0].mlp_s.fc1.weight.mul_(1.5)
    with pytest.raises(RuntimeError, match="modified in place"):
        loss.backward()


def test_no_torch_fallback_unsupported_training_configurations_raise(cuda_device):
    """There is exactly one backward implementation (mb_backward): what it does not cover raises instead of rerouting."""
    torch.manual_seed(0)
    m = DSTformer(dim_in=3, dim_out=17, dim_feat=256, dim_rep=512, depth=1, num_heads=8, mlp_ratio=2,
                  norm_layer=partial(nn.LayerNorm, eps=1e-6)).to(cuda_device).train()
    x = torch.from_numpy(O.make_input(1, 4, 17, 1)).to(cuda_device)
    with pytest.raises(NotImplementedError, match="dim_out"):
        m(x)
    with torch.no_grad():
        assert m(x).shape == (1, 4, 17, 17)             # inference with a wide head is fine


def test_native_backward_without_fusion_head(cuda_device):
    """att_fuse=False (DSTformer.py:350-351: x = (x_st + x_ts) * 0.5) trains through the native backward as well."""
    torch.manual_seed(4)
    m = DSTformer(dim_in=3, dim_out=3, dim_feat=256, dim_rep=512, depth=2, num_heads=8, mlp_ratio=2,
                  norm_layer=partial(nn.LayerNorm, eps=1e-6), att_fuse=False).to(cuda_device).train()
    assert not hasattr(m, "ts_attn")
    B, F = 2, 12
    x = torch.from_numpy(O.make_input(B, F, 17, 3)).to(cuda_device)
    w = torch.randn(B, F, 17, 3, generator=torch.Generator().manual_seed(2)).to(cuda_device)
    out = m(x)
    (out * w).sum().backward()
    live = [p for p in m._ordered_params() if p is not None]
    ps = [p.detach().double().requires_grad_(True) for p in live]
    y = recompute_forward(m, x.double(), False, None, ps)
    assert float((out.detach().double() - y.detach()).abs().max()) < 1e-3 * float(y.abs().max())
    ref = torch.autograd.grad((y * w.double()).sum(), ps)
    for p, gr in zip(live, ref):
        den = float(gr.norm())
        if den > 0:
            assert float((p.grad.double() - gr).norm()) / den < REL_L2


def test_training_step_reduces_loss(cuda_device):
    """Ten AdamW steps on a fixed batch through forward_train/backward/repack: the loss must fall (train.py:149-176)."""
    m = _module(cuda_device, 256, 2, 8, 2, seed=2)
    x = torch.from_numpy(O.make_input(4, 16, 17, 6)).to(cuda_device)
    target = torch.from_numpy(O.make_input(4, 16, 17, 7)).to(cuda_device)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    losses = []
    for _ in range(10):
        opt.zero_grad(set_to_none=True)
        loss = ((m(x) - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)) and losses[-1] < 0.9 * losses[0], losses


def test_backward_error_paths(cuda_device):
    import ctypes

    from motionbert_b200 import _lib
    m = _module(cuda_device, 256, 1, 8, 2)
    x = torch.from_numpy(O.make_input(1, 4, 17, 1)).to(cuda_device)
    out, rep, saved = m._launch_train(x, True)
    lib = _lib.load()
    st = m._state_for(x.device, m.train_math_mode)
    nb = ctypes.c_size_t()
    assert lib.mb_saved_bytes(st.handle, 1, 4, ctypes.byref(nb)) == 0 and nb.value > 0
    assert lib.mb_saved_bytes(st.handle, 1, 1000, ctypes.byref(nb)) < 0
    # too-small saved region is rejected before any launch
    rc = lib.mb_forward_train(st.handle, m._aligned_ptr(st.packed), x.data_ptr(), None, rep.data_ptr(), None,
                              m._aligned_ptr(saved), 1024, m._aligned_ptr(saved), 1 << 30, 1, 4, 0, None)
    assert rc < 0 and b"saved region" in lib.mb_last_error()


def test_frozen_parameters_and_retain_graph(cuda_device):
    """partial_train (lib/utils/learning.py:69-77) freezes subsets of the backbone: frozen tensors must get no .grad,
    the others the same gradient as before; backward(retain_graph=True) may be called twice on one forward."""
    m = _module(cuda_device, 256, 1, 8, 2, seed=6)
    x = torch.from_numpy(O.make_input(2, 8, 17, 4)).to(cuda_device)
    (m(x) ** 2).sum().backward()
    full = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    frozen = [n for n, _ in m.named_parameters() if "attn_t" in n or n == "pos_embed"]
    for n, p in m.named_parameters():
        p.requires_grad_(n not in frozen)
    out = m(x)
    loss = (out ** 2).sum()
    loss.backward(retain_graph=True)
    for n, p in m.named_parameters():
        if n in frozen:
            assert p.grad is None, n
        else:
            assert torch.allclose(p.grad, full[n], rtol=1e-3, atol=1e-5 * float(full[n].abs().max()) + 1e-12), n
    loss.backward()                                   # second pass over the same saved activations: accumulates
    n0 = "blocks_st.0.mlp_s.fc1.weight"
    g = dict(m.named_parameters())[n0].grad
    assert torch.allclose(g, 2 * full[n0], rtol=1e-3, atol=1e-5 * float(full[n0].abs().max()))


def test_native_backward_with_drop_path_and_input_gradient(cuda_device):
    """DropPath (lib/model/drop.py:17-32: per-frame keep mask / keep_prob on every residual branch) and the gradient
    w.r.t. the pose input, both through the native backward, against fp64 autograd with the SAME mask."""
    from motionbert_b200._autograd import DSTformerFunction
    m = _module(cuda_device, 256, 2, 8, 2, seed=8)
    B, F = 3, 10
    x = torch.from_numpy(O.make_input(B, F, 17, 2)).to(cuda_device).requires_grad_(True)
    w = torch.randn(B, F, 17, 3, generator=torch.Generator().manual_seed(3)).to(cuda_device)
    g = torch.Generator().manual_seed(4)
    keep = 0.7
    dp = ((keep + torch.rand(16, B * F, generator=g)).floor() / keep).to(cuda_device).contiguous()
    assert float(dp.min()) == 0.0 and float(dp.max()) > 1.0
    params = m._ordered_params()
    out = DSTformerFunction.apply(m, x, False, dp, *params)
    (out * w).sum().backward()
    ps = [p.detach().double().requires_grad_(True) for p in params]
    xr = x.detach().double().requires_grad_(True)
    y = recompute_forward(m, xr, False, dp.double(), ps)
    assert float((out.detach().double() - y.detach()).abs().max()) < 1e-3 * float(y.abs().max())
    ref = torch.autograd.grad((y * w.double()).sum(), [xr] + ps)
    gx_ref = ref[0]
    rel_x = float((x.grad.double() - gx_ref).norm() / gx_ref.norm())
    print(f"input-gradient rel L2 error {rel_x:.3e}")
    assert rel_x < REL_L2
    _compare(m, list(ref[1:]), "drop_path")


def test_single_pass_mode_saves_attention_operands(cuda_device):
    """set_math_mode('bf16') (the training configuration of bench.py --mode train): the forward keeps the qkv and
    attention-output planes of every attention sublayer in the saved region and the backward uses them instead of
    recomputing.  The forward itself is bf16 single-pass here, so the bar against fp64 autograd is the bf16 forward's."""
    import ctypes

    from motionbert_b200 import _lib
    m = _module(cuda_device, 256, 2, 8, 2, seed=12)
    B, F = 2, 40
    x = torch.from_numpy(O.make_input(B, F, 17, 9)).to(cuda_device)
    w = torch.randn(B, F, 17, 3, generator=torch.Generator().manual_seed(6)).to(cuda_device)
    lib = _lib.load()
    nb3, nb1 = ctypes.c_size_t(), ctypes.c_size_t()
    _lib.check(lib.mb_saved_bytes(m._state_for(x.device, m.train_math_mode).handle, B, F, ctypes.byref(nb3)))
    m.set_math_mode("bf16")
    _lib.check(lib.mb_saved_bytes(m._state_for(x.device, m.train_math_mode).handle, B, F, ctypes.byref(nb1)))
    M = B * F * 17
    assert nb1.value - nb3.value >= 8 * M * 4 * 256 * 2          # 4 attention sublayers x depth 2 x (3C + C) bf16
    out = m(x)
    (out * w).sum().backward()
    grads_ref, y_ref = _reference_grads(m, x, w, False)
    assert float((out.detach().double() - y_ref).abs().max()) < 3e-2 * float(y_ref.abs().max())
    names = [n for n, _ in m.named_parameters()]
    order = {id(p): i for i, p in enumerate(m._ordered_params())}
    worst = (0.0, "")
    for n, p in m.named_parameters():
        if n.startswith("ts_attn."):
            continue
        gr = grads_ref[order[id(p)]]
        den = float(gr.norm())
        if den == 0:
            continue
        rel = float((p.grad.double() - gr).norm()) / den
        cos = float((p.grad.double() * gr).sum()) / (float(p.grad.double().norm()) * den)
        worst = max(worst, (rel, n))
        assert rel < 1e-1 and cos > 0.995, (n, rel, cos)
    print(f"bf16 mode (saved attention operands): worst per-parameter rel L2 {worst[0]:.3e} ({worst[1]}) over {len(names)} tensors")


GRAD_GOLD = ["grads_base_b2_f27", "grads_lite_b2_f27", "grads_base_b1_f243", "grads_lite_b2_f40_rep"]


@pytest.mark.parametrize("name", GRAD_GOLD)
def test_native_backward_matches_reference_gradient_fixtures(cuda_device, name):
    """mb_backward against float64 autograd through the REAL reference module (fixtures of oracle/make_golden_grads.py):
    same perturbed parameters, same clip, loss = sum(y * w) with the fixture's seeded w.  Per tensor: relative L2 error
    over the 512 sampled entries, the norm and the input gradient."""
    from conftest import GOLD, build_module
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.EncoderConfig(dim_feat=int(g["dim_feat"]), mlp_ratio=float(g["mlp_ratio"]))
    B, F, return_rep = int(g["B"]), int(g["F"]), bool(int(g["return_rep"]))
    m = build_module(cfg, O.make_params(cfg, int(g["param_seed"])), cuda_device).train()
    x = torch.from_numpy(O.make_input(B, F, cfg.num_joints, int(g["input_seed"]))).to(cuda_device).requires_grad_(True)
    y = m.get_representation(x) if return_rep else m(x)
    w = torch.from_numpy(O.fixture_out_weight(tuple(y.shape), int(g["w_seed"]))).float().to(cuda_device)
    loss = (y * w).sum()
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 2e-3 * abs(float(g["loss"])) + 1e-2
    dx_ref = torch.from_numpy(g["dx"]).to(cuda_device)
    rel_x = float((x.grad.double() - dx_ref).norm() / dx_ref.norm())
    worst, gate = (0.0, ""), {}
    params = m._ordered_params()
    names = list(O.param_shapes(cfg).keys())
    for i, (n, p) in enumerate(zip(names, params)):
        ref = torch.from_numpy(g[f"val_{i}"]).to(cuda_device)
        if float(g[f"norm_{i}"]) == 0.0:                      # head.* on the representation path
            assert p.grad is None, n
            continue
        got = p.grad.double().reshape(-1)[torch.from_numpy(g[f"idx_{i}"]).to(cuda_device)]
        if n.startswith("ts_attn."):                          # judged jointly (weight + bias), see _compare
            gate.setdefault(n.rsplit(".", 1)[0], []).append((got, ref))
            continue
        rel = float((got - ref).norm() / ref.norm().clamp_min(1e-300))
        nrm = float(p.grad.double().norm()) / float(g[f"norm_{i}"])
        worst = max(worst, (rel, n))
        assert rel < REL_L2 and abs(nrm - 1.0) < REL_L2, (n, rel, nrm)
    for n, lst in gate.items():
        got, ref = torch.cat([a for a, _ in lst]), torch.cat([b for _, b in lst])
        assert float((got - ref).norm() / ref.norm()) < JOINT_GATE_REL, n
    print(f"[{name}] vs reference autograd: worst sampled rel L2 {worst[0]:.3e} ({worst[1]}), input gradient {rel_x:.3e}")
    assert rel_x < REL_L2


def test_native_backward_full_depth_full_length_batch8(cuda_device):
    """Depth-5 DSTformer-base at the BASELINE sequence length (T = 243) and B = 8 (M = 33,048 tokens: 130 GEMM row tiles,
    16 temporal query tiles per joint-head): native backward vs fp64 autograd of the restatement on the same device."""
    m = _module(cuda_device, 512, 5, 8, 2, seed=21)
    B, F = 8, 243
    x = torch.from_numpy(O.make_input(B, F, 17, 31)).to(cuda_device)
    w = torch.randn(B, F, 17, 3, generator=torch.Generator().manual_seed(9)).to(cuda_device)
    out = m(x)
    (out * w).sum().backward()
    torch.cuda.synchronize(cuda_device)
    grads_ref, y_ref = _reference_grads(m, x, w, False)
    assert float((out.detach().double() - y_ref).abs().max()) < 1e-3 * float(y_ref.abs().max())
    _compare(m, grads_ref, "base_full_f243_b8")
