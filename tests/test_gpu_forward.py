"""Whole-path parity on the GPU through the product class (-> ctypes -> C ABI -> sm_100a kernels):
against the committed reference outputs (tests/golden), against the CPU oracle on fresh seeded inputs, and
through size-independent properties at BASELINE's full sequence length."""
import numpy as np
import pytest
import torch

from conftest import build_module, golden_names, load_case, rel_token_err
from motionbert_b200 import _lib
from oracle import dstformer_oracle as O

pytestmark = pytest.mark.gpu

# north-star tolerances (BASELINE.json): per-token relative 1e-3 in fp32 mode; MPJPE within 0.1 mm.
# The golden skeletons are unit-scale (|out| ~ 0.4); H36M poses are normalised by ~500 mm per unit
# (train.py evaluate), so 0.1 mm == 2e-4 units.
TOK_REL = 1e-3
MPJPE_UNITS = 2e-4


def _run(m, x, dev):
    with torch.no_grad():
        xt = torch.from_numpy(x).to(dev)
        out = m(xt)
        rep = m.get_representation(xt)
    torch.cuda.synchronize(dev)
    return out.cpu().numpy(), rep.cpu().numpy()


def _check_against_golden(out, rep, g, cfg, label):
    M = out.shape[0] * out.shape[1] * cfg.num_joints
    assert np.isfinite(out).all() and np.isfinite(rep).all(), label
    mean_o, max_o = rel_token_err(out, g["out"])
    mean_r, max_r = rel_token_err(rep.reshape(M, -1)[g["rep_idx"]], g["rep_rows"])
    mp = O.mpjpe(out.astype(np.float64), g["out64"])
    mp_ref = O.mpjpe(g["out"].astype(np.float64), g["out64"])
    print(f"[{label}] out rel mean/max {mean_o:.2e}/{max_o:.2e}  rep rel {mean_r:.2e}/{max_r:.2e}  "
          f"MPJPE vs fp64 truth {mp:.2e} (reference fp32: {mp_ref:.2e})")
    assert max_r < TOK_REL and mean_o < TOK_REL, label
    assert abs(mp - mp_ref) < MPJPE_UNITS and mp < MPJPE_UNITS, label


@pytest.mark.parametrize("name", golden_names())
def test_forward_matches_reference_golden_simt_reference_kernels(cuda_device, name):
    """Bring-up path: CUDA-core GEMM/attention reference kernels (isolates the non-tensor-core code)."""
    cfg, P, x, g = load_case(name)
    if x.shape[0] * x.shape[1] > 64:
        pytest.skip("CUDA-core reference GEMM is only run on the small cases")
    m = build_module(cfg, P, cuda_device).use_test_library().set_math_mode("bf16x3")   # CUDA-core kernels: test twin, bf16 planes
    m._kernel_flags = _lib.MB_FLAG_REF_GEMM | _lib.MB_FLAG_REF_ATTN_T | _lib.MB_FLAG_REF_ATTN_S
    out, rep = _run(m, x, cuda_device)
    _check_against_golden(out, rep, g, cfg, name + "/simt")


@pytest.mark.parametrize("mode", ["f16c", "bf16x3"])
@pytest.mark.parametrize("name", golden_names())
def test_forward_matches_reference_golden(cuda_device, name, mode):
    """The product path (tcgen05 GEMMs + tcgen05 attention) in both fp32-parity arithmetic modes: F16C (fp16 pass + one
    e5m2 compensation pass; the inference default) and BF16x3 (three bf16 passes; the training forward)."""
    cfg, P, x, g = load_case(name)
    m = build_module(cfg, P, cuda_device).set_math_mode(mode)
    out, rep = _run(m, x, cuda_device)
    _check_against_golden(out, rep, g, cfg, name + "/" + mode)


def test_product_library_rejects_the_test_kernel_flags(cuda_device):
    """The CUDA-core / first-generation kernels are not in libmotionbert_b200.so: their flags fail loudly, no reroute."""
    cfg, P, x, g = load_case("lite_b2_f27")
    m = build_module(cfg, P, cuda_device).set_math_mode("bf16x3")
    m._kernel_flags = _lib.MB_FLAG_REF_GEMM
    with pytest.raises(_lib.MbError, match="test"):
        _run(m, x, cuda_device)
    m._kernel_flags = _lib.MB_FLAG_REF_ATTN_T
    with pytest.raises(_lib.MbError, match="test"):
        _run(m, x, cuda_device)


def test_f16c_attention_equals_bf16x3_attention_within_tolerance(cuda_device):
    """A/B inside the F16C mode: the F16C attention kernels against the BF16x3 attention kernels fed with bf16 planes."""
    cfg, P, x, g = load_case("base_b2_f130")
    m = build_module(cfg, P, cuda_device)
    o1, r1 = _run(m, x, cuda_device)
    m._kernel_flags = _lib.MB_FLAG_ATTN_BF16X3
    o2, r2 = _run(m, x, cuda_device)
    assert rel_token_err(r1, r2)[1] < 5e-4
    _check_against_golden(o2, r2, g, cfg, "base_b2_f130/f16c+bf16x3-attention")


def test_default_math_mode_is_f16c_for_inference_and_bf16x3_for_gradients(cuda_device):
    cfg, P, x, g = load_case("lite_b2_f27")
    m = build_module(cfg, P, cuda_device)
    assert m.math_mode == _lib.MB_MATH_F16C and m.train_math_mode == _lib.MB_MATH_BF16X3
    _run(m, x, cuda_device)
    assert (cuda_device.index or 0, _lib.MB_MATH_F16C) in m._dev_state


def test_forward_first_generation_1cta_gemm_still_matches(cuda_device):
    cfg, P, x, g = load_case("base_b2_f27")
    m = build_module(cfg, P, cuda_device).use_test_library().set_math_mode("bf16x3")
    m._kernel_flags = _lib.MB_FLAG_GEMM_1CTA
    out, rep = _run(m, x, cuda_device)
    _check_against_golden(out, rep, g, cfg, "base_b2_f27/1cta")


def test_forward_matches_cpu_oracle_on_fresh_inputs(cuda_device):
    cfg = O.LITE
    P = O.make_params(cfg, 99)
    x = O.make_input(3, 20, cfg.num_joints, 123)
    o_ref, r_ref = O.forward(P, x, cfg, np.float64)
    m = build_module(cfg, P, cuda_device)
    out, rep = _run(m, x, cuda_device)
    assert rel_token_err(rep, r_ref)[1] < TOK_REL
    assert O.mpjpe(out.astype(np.float64), o_ref) < MPJPE_UNITS


def test_reference_init_known_answer(cuda_device):
    """torch.manual_seed(0) reference init + the survey's probe input (SURVEY.md 8c known-answer values)."""
    from conftest import manifest
    kat = manifest()["init_kat"]["base"]
    torch.manual_seed(0)
    m = build_module(O.BASE, None, cuda_device)
    x = torch.rand(2, 27, 17, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1
    with torch.no_grad():
        out = m(x.to(cuda_device)).cpu()
        rep = m.get_representation(x.to(cuda_device)).cpu()
    assert abs(float(out.double().sum()) - kat["out_sum"]) < 2e-2
    np.testing.assert_allclose(out[0, 0, 0].numpy(), kat["out_000"], atol=2e-5)
    np.testing.assert_allclose(rep[1, 26, 16, :3].numpy(), kat["rep_last3"], atol=5e-5)
    assert abs(float(rep.double().sum()) - kat["rep_sum"]) < 0.5


def test_full_length_properties_batch_independence_and_determinism(cuda_device):
    """BASELINE config-2 sequence shape (T=243, base) at a batch the test box holds: sequences are independent
    units, so a sequence's output must not depend on its batch mates; and the path is
    deterministic (bitwise) run to run."""
    cfg = O.BASE
    P = O.make_params(cfg, 11)
    m = build_module(cfg, P, cuda_device)
    x = O.make_input(6, 243, cfg.num_joints, 77)
    out, rep = _run(m, x, cuda_device)
    out2, rep2 = _run(m, x, cuda_device)
    assert np.array_equal(out, out2) and np.array_equal(rep, rep2)
    o1, r1 = _run(m, x[4:5], cuda_device)
    assert rel_token_err(r1[0], rep[4])[1] < 2e-5
    # and it still matches the committed reference output for the sequence that has a golden (base_b1_f243)
    cfg2, P2, x2, g = load_case("base_b1_f243")
    xx = np.concatenate([x2, x[:3]], axis=0)
    oo, _ = _run(m, xx, cuda_device)
    assert O.mpjpe(oo[0:1].astype(np.float64), g["out64"]) < MPJPE_UNITS


def test_bf16_single_pass_mode_is_within_bf16_tolerance(cuda_device):
    cfg, P, x, g = load_case("base_b2_f27")
    m = build_module(cfg, P, cuda_device).set_math_mode("bf16")
    out, rep = _run(m, x, cuda_device)
    mean_o, max_o = rel_token_err(out, g["out"])
    print(f"[bf16 1-pass] out rel mean/max {mean_o:.2e}/{max_o:.2e}")
    assert mean_o < 5e-2


def test_drop_path_training_matches_torch_recompute(cuda_device):
    from functools import partial

    import torch.nn as nn

    from motionbert_b200 import DSTformer
    from oracle.dstformer_torch_autograd import recompute_forward
    cfg = O.LITE
    torch.manual_seed(3)
    m = DSTformer(dim_in=3, dim_out=3, dim_feat=256, dim_rep=512, depth=5, num_heads=8, mlp_ratio=4,
                  norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_path_rate=0.3).to(cuda_device).train()
    P = O.make_params(cfg, 4)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()})
    x = torch.from_numpy(O.make_input(2, 9, 17, 5)).to(cuda_device)
    dp = m._drop_path_scale(2, 9, x.device)
    assert dp is not None and dp.shape == (40, 18) and float(dp.min()) == 0.0
    with torch.no_grad():
        exp = recompute_forward(m, x.double(), False, dp.double(), [p.double() for p in m._ordered_params()])
        for mode in ("bf16x3", "f16c"):                      # DropPath lives in the residual epilogue of every arithmetic mode
            m.set_math_mode(mode)
            out, _ = m._launch(x, True, False, dp)
            mean_e, max_e = rel_token_err(out.cpu().numpy(), exp.cpu().numpy())
            print(f"[drop_path/{mode}] out per-token rel mean {mean_e:.2e} max {max_e:.2e}")
            assert mean_e < TOK_REL and (mode != "bf16x3" or max_e < TOK_REL)


def test_autograd_backward_runs_and_matches_torch(cuda_device):
    cfg = O.LITE
    P = O.make_params(cfg, 8)
    m = build_module(cfg, P, cuda_device).train()
    x = torch.from_numpy(O.make_input(2, 6, 17, 2)).to(cuda_device)
    out = m(x)
    loss = (out ** 2).mean()
    loss.backward()
    g = m.head.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    assert m.blocks_st[0].attn_t.qkv.weight.grad is not None


def test_host_buffer_entry_point(cuda_device):
    """mb_forward_host: the C-ABI call with HOST buffers (what bench.py's e2e leg times)."""
    import ctypes
    cfg, P, x, g = load_case("lite_b2_f27")
    m = build_module(cfg, P, cuda_device)
    _run(m, x, cuda_device)                     # creates handle + packs weights
    st = m._state_for(cuda_device)
    lib = _lib.load()
    B, F = x.shape[:2]
    nb = ctypes.c_size_t()
    _lib.check(lib.mb_workspace_bytes_host(st.handle, B, F, 1, 0, ctypes.byref(nb)))
    ws = torch.empty(nb.value + 1024, dtype=torch.uint8, device=cuda_device)
    xh = torch.from_numpy(x).pin_memory()
    oh = torch.empty(B, F, 17, 3).pin_memory()
    with torch.cuda.device(cuda_device):
        _lib.check(lib.mb_forward_host(st.handle, m._aligned_ptr(st.packed), xh.data_ptr(), oh.data_ptr(), None,
                                       m._aligned_ptr(ws), nb.value, B, F, 0,
                                       torch.cuda.current_stream().cuda_stream), "mb_forward_host")
    assert O.mpjpe(oh.numpy().astype(np.float64), g["out64"]) < MPJPE_UNITS


def test_error_paths(cuda_device):
    import ctypes
    cfg, P, x, g = load_case("lite_b2_f27")
    m = build_module(cfg, P, cuda_device)
    _run(m, x, cuda_device)
    st = m._state_for(cuda_device)
    lib = _lib.load()
    xt = torch.from_numpy(x).to(cuda_device)
    out = torch.empty(2, 27, 17, 3, device=cuda_device)
    ws = torch.empty(4096, dtype=torch.uint8, device=cuda_device)
    rc = lib.mb_forward(st.handle, m._aligned_ptr(st.packed), xt.data_ptr(), out.data_ptr(), None, None,
                        m._aligned_ptr(ws), 1024, 2, 27, 0, None)
    assert rc == -6 and b"workspace" in lib.mb_last_error()
    rc = lib.mb_forward(st.handle, m._aligned_ptr(st.packed), xt.data_ptr(), out.data_ptr(), None, None,
                        m._aligned_ptr(ws), 1024, 2, 300, 0, None)
    assert rc == -1 and b"maxlen" in lib.mb_last_error()
    rc = lib.mb_forward(st.handle, m._aligned_ptr(st.packed), None, out.data_ptr(), None, None,
                        m._aligned_ptr(ws), 1024, 2, 27, 0, None)
    assert rc == -2


def test_cuda_graph_capture_and_replay(cuda_device):
    """The whole forward is capturable (no allocation / sync inside the library) and replays bit-identically."""
    cfg, P, x, g = load_case("lite_b2_f27")
    m = build_module(cfg, P, cuda_device)
    ref_out, _ = _run(m, x, cuda_device)
    run = m.make_graphed(2, 27)
    xt = torch.from_numpy(x).to(cuda_device)
    o1 = run(xt).clone()
    o2 = run(xt * 0.5).clone()
    o3 = run(xt).clone()
    torch.cuda.synchronize()
    assert np.array_equal(o1.cpu().numpy(), ref_out) and torch.equal(o1, o3) and not torch.equal(o1, o2)


def test_data_parallel_wrapper_two_gpus():
    """The reference wraps the backbone in nn.DataParallel (train.py:256-258): replicas on 2 devices, one Python
    thread each, must reproduce the single-device result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg, P, x, g = load_case("lite_b5_f30")
    dev0 = torch.device("cuda:0")
    m = build_module(cfg, P, dev0)
    single, _ = _run(m, x, dev0)
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    with torch.no_grad():
        out = dp(torch.from_numpy(x).to(dev0)).cpu().numpy()
        out2 = dp(torch.from_numpy(x).to(dev0)).cpu().numpy()
    assert out.shape == single.shape
    assert rel_token_err(out, single)[1] < 2e-5 and np.array_equal(out, out2)
    assert O.mpjpe(out.astype(np.float64), g["out64"]) < MPJPE_UNITS


def test_data_parallel_training_backward_two_gpus():
    """train.py:256-258 + :205: the backbone wrapped in nn.DataParallel must TRAIN -- replicas keep their parameters as
    plain attributes (`_parameters` is empty), so the needs-gradient decision goes through `_ordered_params()`; the
    gradients reduced onto device 0 must equal the single-device gradients of the same global batch."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cfg, P, x, g = load_case("lite_b5_f30")
    dev0 = torch.device("cuda:0")
    xt = torch.from_numpy(x[:4]).to(dev0)
    w = torch.randn(4, 30, 17, 3, generator=torch.Generator().manual_seed(3)).to(dev0)
    m = build_module(cfg, P, dev0).train()
    (m(xt) * w).sum().backward()
    single = {n: p.grad.clone() for n, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    dp = torch.nn.DataParallel(m, device_ids=[0, 1])
    out = dp(xt)
    assert out.requires_grad
    (out * w).sum().backward()
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        den = float(single[n].norm())
        if den > 0 and not n.startswith("ts_attn."):
            assert float((p.grad - single[n]).norm()) / den < 2e-2, n      # bf16 backward, different batch split


def test_input_variants_noncontiguous_strided_and_double(cuda_device):
    """Callers hand over slices / permuted views (train.py:160-172 builds batch_input on the fly); the module must
    accept any strided float tensor and produce a fresh contiguous writable result."""
    cfg, P, x, g = load_case("lite_b2_f27")
    m = build_module(cfg, P, cuda_device)
    ref, _ = _run(m, x, cuda_device)
    xt = torch.from_numpy(x).to(cuda_device)
    big = torch.zeros(2, 27, 17, 5, device=cuda_device)
    big[..., 1:4] = xt
    with torch.no_grad():
        o1 = m(big[..., 1:4])                                   # non-contiguous last-dim slice
        o2 = m(xt.permute(1, 0, 2, 3).contiguous().permute(1, 0, 2, 3))   # permuted strides
        o3 = m(xt.double())                                     # wrong dtype is converted, not rejected
    for o in (o1, o2, o3):
        assert o.is_contiguous() and o.dtype == torch.float32
        assert np.array_equal(o.cpu().numpy(), ref)
    o1 += 1.0                                                   # callers edit the result in place (infer_wild.py:82-87)


def test_two_streams_and_odd_batch_tail(cuda_device):
    """Work is enqueued on the caller's current stream; odd batches leave a partial 256-row tile at the end."""
    cfg, P, x, g = load_case("base_b3_f16")
    m = build_module(cfg, P, cuda_device)
    ref, _ = _run(m, x, cuda_device)
    xt = torch.from_numpy(x).to(cuda_device)
    s1 = torch.cuda.Stream(device=cuda_device)
    torch.cuda.synchronize()
    with torch.no_grad(), torch.cuda.stream(s1):
        o = m(xt)
    s1.synchronize()
    assert np.array_equal(o.cpu().numpy(), ref)
    with torch.no_grad():
        o5 = m(torch.cat([xt, xt[:2]], 0))                      # B = 5
    assert np.array_equal(o5[:3].cpu().numpy(), ref) and np.array_equal(o5[3:].cpu().numpy(), ref[:2])


def test_small_shapes_are_graphed_automatically_and_track_parameter_updates(cuda_device):
    """Launch-bound inference shapes: from the third call on the forward is a CUDA-graph replay (one launch instead of
    ~90) with bit-identical results; a parameter update drops the captured graph (no stale weights)."""
    cfg, P, x, g = load_case("lite_b2_f27")
    m = build_module(cfg, P, cuda_device)
    xt = torch.from_numpy(x).to(cuda_device)
    m.auto_graph_max_tokens = 0
    with torch.no_grad():
        ref = m(xt).clone()
        m.auto_graph_max_tokens = 16384
        outs = [m(xt) for _ in range(5)]
        st = m._state_for(cuda_device)
        ent = st.graphs[(2, 27, False, 0)]
        assert ent["graph"] is not None and (2, 27) in st.pinned
        assert all(torch.equal(o, ref) for o in outs) and outs[3].data_ptr() != outs[4].data_ptr()
        x2 = torch.from_numpy(O.make_input(2, 27, cfg.num_joints, 77)).to(cuda_device)
        m.auto_graph_max_tokens = 0
        ref2 = m(x2).clone()
        m.auto_graph_max_tokens = 16384
        assert torch.equal(m(x2), ref2)                                   # replay on new input data
        m.head.bias.add_(0.25)                                            # in-place update: _version changes
        o_new = m(xt)
        assert torch.allclose(o_new, ref + 0.25, atol=1e-6) and not torch.equal(o_new, ref)
        for _ in range(3):
            o_new2 = m(xt)                                                # re-captured with the new weights
        assert torch.equal(o_new2, o_new) and st.graphs[(2, 27, False, 0)]["graph"] is not None
        big = torch.from_numpy(O.make_input(8, 243, cfg.num_joints, 5)).to(cuda_device)   # 33048 tokens: never graphed
        for _ in range(4):
            m(big)
        assert st.graphs.get((8, 243, False, 0)) is None
