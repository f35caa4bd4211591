"""The fused MLP sublayer kernel (mlp_fused.cuh: fc1 -> GELU -> fc2 -> residual in ONE launch, hidden activation kept in an
L2-resident ring) against the two-GEMM form it replaces (reference: lib/model/DSTformer.py:79-85, :242-249).

Both forms run the same MMAs in the same order per K block and the same epilogue arithmetic, so the comparison bar is
BIT EQUALITY of the whole network's outputs -- any race on the hidden ring, a missed dependency flag or a wrong ring slot
shows up as a difference.  Parity of the fused default against the reference itself is test_gpu_forward.py (goldens)."""
import numpy as np
import pytest
import torch

from conftest import build_module, load_case
from motionbert_b200 import _lib
from oracle import dstformer_oracle as O

pytestmark = pytest.mark.gpu


def _run(m, xt):
    with torch.no_grad():
        out = m(xt)
        rep = m.get_representation(xt)
    torch.cuda.synchronize(xt.device)
    return out.clone(), rep.clone()


def _handle(m, dev):
    return m._state_for(dev).handle


VARIANTS = {                                    # kernel-flag sets of the fused kernel (default 0: block order, L2 ring, evict_last)
    "ring": 0,
    "ring-nohint": _lib.MB_FLAG_MLP_NO_HINT,
    "noring": _lib.MB_FLAG_MLP_NO_RING,
    "noring-nohint": _lib.MB_FLAG_MLP_NO_RING | _lib.MB_FLAG_MLP_NO_HINT,
}


@pytest.mark.parametrize("name", ["lite_b2_f27", "lite_b1_f243", "base_b1_f1", "base_b3_f16", "base_b2_f130", "base_b1_f243"])
def test_fused_mlp_is_bit_identical_to_the_two_gemm_form(cuda_device, name):
    cfg, P, x, g = load_case(name)
    m = build_module(cfg, P, cuda_device)
    xt = torch.from_numpy(x).to(cuda_device)
    m._kernel_flags = _lib.MB_FLAG_MLP_SPLIT
    o_s, r_s = _run(m, xt)                              # fc1 GEMM + fc2 GEMM
    assert torch.isfinite(r_s).all()
    for label, fl in VARIANTS.items():
        m._kernel_flags = fl
        o_f, r_f = _run(m, xt)
        assert torch.equal(r_f, r_s) and torch.equal(o_f, o_s), (name, label)


def test_fused_mlp_launch_count(cuda_device):
    cfg, P, x, g = load_case("lite_b2_f27")
    m = build_module(cfg, P, cuda_device)
    _run(m, torch.from_numpy(x).to(cuda_device))
    lib = _lib.load()
    h = _handle(m, cuda_device)
    fused = _lib.check(lib.mb_forward_launch_count(h, 1, 0))
    split = _lib.check(lib.mb_forward_launch_count(h, 1, _lib.MB_FLAG_MLP_SPLIT))
    assert split - fused == 4 * cfg.depth and fused == 1 + cfg.depth * 17 + 2


@pytest.mark.parametrize("model,B,F", [("base", 6, 243), ("lite", 24, 81), ("base", 7, 200)])
def test_fused_mlp_many_token_blocks_per_cta_pair(cuda_device, model, B, F):
    """More 256-row token blocks than CTA pairs: every pair walks several blocks, the ring slot is rewritten, the ragged
    last block is clipped.  Bit equality with the two-GEMM form, and run-to-run determinism of the fused form."""
    cfg = O.BASE if model == "base" else O.LITE
    P = O.make_params(cfg, 7)
    x = O.make_input(B, F, cfg.num_joints, 11)
    assert (B * F * cfg.num_joints + 255) // 256 > 74
    m = build_module(cfg, P, cuda_device)
    xt = torch.from_numpy(x).to(cuda_device)
    m._kernel_flags = _lib.MB_FLAG_MLP_SPLIT
    o_s, r_s = _run(m, xt)
    assert torch.isfinite(r_s).all()
    for label, fl in VARIANTS.items():
        m._kernel_flags = fl
        o_f, r_f = _run(m, xt)
        o_f2, r_f2 = _run(m, xt)
        assert torch.equal(r_f, r_f2) and torch.equal(o_f, o_f2), label
        assert torch.equal(r_f, r_s) and torch.equal(o_f, o_s), label


@pytest.mark.parametrize("dim,ratio,B,F", [(512, 4, 3, 50), (256, 1, 3, 50), (256, 2, 2, 33), (512, 1, 2, 40), (512, 4, 8, 243)])
def test_fused_mlp_tile_geometries_no_shipped_config_uses(cuda_device, dim, ratio, B, F):
    """hidden = 2048 (8 fc1 tiles per token block, the kernel's maximum), hidden = C (one fc1 tile), C = hidden = 512 (2 + 2):
    bit equality with the two-GEMM form, and parity with the float64 oracle (profiles/r02m_mlp_geometry_check.log)."""
    cfg = O.EncoderConfig(dim_feat=dim, mlp_ratio=ratio, depth=2)
    P = O.make_params(cfg, 5)
    xn = O.make_input(B, F, cfg.num_joints, 9)
    xt = torch.from_numpy(xn).to(cuda_device)
    m = build_module(cfg, P, cuda_device)
    with torch.no_grad():
        m._kernel_flags = _lib.MB_FLAG_MLP_SPLIT
        r_s = m.get_representation(xt).clone()
        m._kernel_flags = 0
        r_f = m.get_representation(xt).clone()
        m._kernel_flags = _lib.MB_FLAG_MLP_NO_RING
        r_n = m.get_representation(xt).clone()
    assert torch.isfinite(r_f).all() and torch.equal(r_f, r_s) and torch.equal(r_n, r_s)
    if B * F < 400:
        _o, r_ref = O.forward(P, xn[:1], cfg, np.float64)
        d = r_f[:1].cpu().numpy().astype(np.float64) - r_ref
        assert float((np.linalg.norm(d, axis=-1) / np.linalg.norm(r_ref, axis=-1)).max()) < 1e-3
