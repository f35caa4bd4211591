"""CPU oracle for the DSTformer encoder hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch numpy restatement of the reference algorithm
(`/root/reference/lib/model/DSTformer.py`); every function cites the reference
file:line it follows.  It is *not* part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import it, and only as the checker / reported CPU baseline.  The product
path (`motionbert_b200`) never imports anything from `oracle/` and fails loudly
when the CUDA library is missing.

Parity pinning: the reference has no tests / golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against the reference module
itself, imported in the build container by `oracle/make_golden.py`
(max |diff| recorded in `tests/golden/MANIFEST.json`), and against the
known-answer values recorded in SURVEY.md section 8(c).

All arithmetic is plain floating point in `dtype` (float32 to mirror the
reference, float64 for a "truth" run).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

try:  # scipy is in the image; fall back to math.erf vectorised if absent
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


@dataclass(frozen=True)
class EncoderConfig:
    """Constructor arguments of `DSTformer.__init__` (DSTformer.py:270-273) as
    the factory passes them (`lib/utils/learning.py:83-85`)."""
    dim_in: int = 3
    dim_out: int = 3
    dim_feat: int = 512
    dim_rep: int = 512
    depth: int = 5
    num_heads: int = 8
    mlp_ratio: float = 2
    num_joints: int = 17
    maxlen: int = 243
    eps: float = 1e-6          # learning.py:84  partial(nn.LayerNorm, eps=1e-6)

    @property
    def hidden(self) -> int:
        return int(self.dim_feat * self.mlp_ratio)      # DSTformer.py:232

    @property
    def head_dim(self) -> int:
        return self.dim_feat // self.num_heads           # DSTformer.py:92


BASE = EncoderConfig(dim_feat=512, mlp_ratio=2)     # configs/pretrain/MB_pretrain.yaml:18-24
LITE = EncoderConfig(dim_feat=256, mlp_ratio=4)     # configs/pretrain/MB_lite.yaml:18-24


# --------------------------------------------------------------------------- #
# parameter tree (DSTformer.py:270-311) -- names/shapes of the 260 tensors
# --------------------------------------------------------------------------- #
def param_shapes(cfg: EncoderConfig) -> "dict[str, tuple]":
    """state_dict names and shapes in the reference registration order."""
    C, hid, J = cfg.dim_feat, cfg.hidden, cfg.num_joints
    d: dict[str, tuple] = {}
    d["temp_embed"] = (1, cfg.maxlen, 1, C)             # :301
    d["pos_embed"] = (1, J, C)                           # :302
    d["joints_embed.weight"] = (C, cfg.dim_in)           # :276
    d["joints_embed.bias"] = (C,)
    for stream in ("blocks_st", "blocks_ts"):            # :280-291
        for i in range(cfg.depth):
            p = f"{stream}.{i}."
            for n in ("norm1_s", "norm1_t"):             # :221-222
                d[p + n + ".weight"] = (C,)
                d[p + n + ".bias"] = (C,)
            for a in ("attn_s", "attn_t"):               # :223-226 (proj built before qkv, :97,103)
                d[p + a + ".proj.weight"] = (C, C)
                d[p + a + ".proj.bias"] = (C,)
                d[p + a + ".qkv.weight"] = (3 * C, C)
                d[p + a + ".qkv.bias"] = (3 * C,)
            for n in ("norm2_s", "norm2_t"):             # :230-231
                d[p + n + ".weight"] = (C,)
                d[p + n + ".bias"] = (C,)
            for m in ("mlp_s", "mlp_t"):                 # :234-235
                d[p + m + ".fc1.weight"] = (hid, C)
                d[p + m + ".fc1.bias"] = (hid,)
                d[p + m + ".fc2.weight"] = (C, hid)
                d[p + m + ".fc2.bias"] = (C,)
    d["norm.weight"] = (C,)                              # :292
    d["norm.bias"] = (C,)
    d["pre_logits.fc.weight"] = (cfg.dim_rep, C)         # :294-297
    d["pre_logits.fc.bias"] = (cfg.dim_rep,)
    d["head.weight"] = (cfg.dim_out, cfg.dim_rep)        # :300
    d["head.bias"] = (cfg.dim_out,)
    for i in range(cfg.depth):                           # :306-311
        d[f"ts_attn.{i}.weight"] = (2, 2 * C)
        d[f"ts_attn.{i}.bias"] = (2,)
    return d


def make_params(cfg: EncoderConfig, seed: int, scale: float = 1.0) -> "dict[str, np.ndarray]":
    """Deterministic *perturbed* parameters for parity tests (the
    reference init leaves ts_attn / LayerNorm affine trivial, so fusion or
    affine bugs would pass unnoticed).  numpy PCG64 -> identical on every box
    with the same numpy.  Not the reference init (that is `DSTformer.__init__`
    of the product class, checked separately against SURVEY's checksums)."""
    rng = np.random.default_rng(seed)
    out: dict[str, np.ndarray] = {}
    for name, shape in param_shapes(cfg).items():
        g = rng.standard_normal(shape, dtype=np.float64)
        if name in ("temp_embed", "pos_embed"):
            v = 0.1 * g
        elif name.startswith("ts_attn") and name.endswith("weight"):
            v = 0.05 * g
        elif name.startswith("ts_attn") and name.endswith("bias"):
            v = 0.5 + 0.3 * g
        elif ".norm" in name or name.startswith("norm."):
            v = (1.0 + 0.2 * g) if name.endswith("weight") else 0.1 * g
        elif name == "joints_embed.weight":
            v = 0.5 * g
        elif name.endswith("weight"):
            fan_in = shape[-1]
            v = scale * g * (1.0 / math.sqrt(fan_in))
        else:  # linear bias
            v = 0.05 * g
        out[name] = v.astype(np.float32)
    return out


def make_input(B: int, F: int, J: int = 17, seed: int = 1) -> np.ndarray:
    """Synthetic 2D skeleton clip: x,y ~ U(-1,1) (crop_scale's
    range, lib/utils/utils_data.py:26-28), confidence ~ U(0,1)."""
    rng = np.random.default_rng(seed)
    x = rng.random((B, F, J, 3), dtype=np.float64)
    x[..., :2] = x[..., :2] * 2.0 - 1.0
    return x.astype(np.float32)


def fixture_out_weight(shape, seed: int) -> np.ndarray:
    """The seeded weight w of the gradient fixtures' scalar loss sum(y * w) (oracle/make_golden_grads.py)."""
    return np.random.default_rng(seed).standard_normal(shape)


def fixture_sample_idx(n: int, nsample: int = 512) -> np.ndarray:
    """Evenly spaced flat indices at which the gradient fixtures keep entries of an n-element tensor."""
    return np.unique(np.linspace(0, n - 1, num=min(nsample, n)).round().astype(np.int64))


# --------------------------------------------------------------------------- #
# forward pieces
# --------------------------------------------------------------------------- #
def _linear(x, w, b):
    """nn.Linear: y = x W^T + b."""
    return x @ w.T + b


def _layer_norm(x, g, b, eps):
    """nn.LayerNorm over the last dim, biased variance (DSTformer.py:221-231,292)."""
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return xc / np.sqrt(var + x.dtype.type(eps)) * g + b


def _gelu(x):
    """nn.GELU() default = exact erf form (DSTformer.py:70,81)."""
    return (0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def _softmax(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


def _mlp(x, P, p):
    """MLP.forward (DSTformer.py:79-85), dropout p=0."""
    h = _gelu(_linear(x, P[p + ".fc1.weight"], P[p + ".fc1.bias"]))
    return _linear(h, P[p + ".fc2.weight"], P[p + ".fc2.bias"])


def _attention(x, P, p, mode, F, cfg):
    """Attention.forward with st_mode 'spatial' (:143-146,178-186) or
    'temporal' (:139-142,188-200).  x: (B*F, J, C)."""
    BF, J, C = x.shape
    H, d = cfg.num_heads, cfg.head_dim
    scale = x.dtype.type(d ** -0.5)                                        # :94
    qkv = _linear(x, P[p + ".qkv.weight"], P[p + ".qkv.bias"])             # (BF,J,3C)
    qkv = qkv.reshape(BF, J, 3, H, d).transpose(2, 0, 3, 1, 4)             # (3,BF,H,J,d) :143
    q, k, v = qkv[0], qkv[1], qkv[2]
    if mode == "spatial":
        att = _softmax((q @ k.transpose(0, 1, 3, 2)) * scale)              # :180-181
        o = att @ v                                                        # (BF,H,J,d) :184
        o = o.transpose(0, 2, 1, 3).reshape(BF, J, C)                      # :185
    else:
        B = BF // F
        qt = q.reshape(B, F, H, J, d).transpose(0, 2, 3, 1, 4)             # (B,H,J,F,d) :190
        kt = k.reshape(B, F, H, J, d).transpose(0, 2, 3, 1, 4)
        vt = v.reshape(B, F, H, J, d).transpose(0, 2, 3, 1, 4)
        att = _softmax((qt @ kt.transpose(0, 1, 2, 4, 3)) * scale)         # :194-195
        o = att @ vt                                                       # (B,H,J,F,d) :198
        o = o.transpose(0, 3, 2, 1, 4).reshape(BF, J, C)                   # :199
    return _linear(o, P[p + ".proj.weight"], P[p + ".proj.bias"])          # :148


def _block(x, P, p, order, F, cfg):
    """Block.forward, 'stage_st' (:240-244) or 'stage_ts' (:245-249); drop_path=0."""
    def ln(t, n):
        return _layer_norm(t, P[p + n + ".weight"], P[p + n + ".bias"], cfg.eps)
    for which in order:
        s = "s" if which == "S" else "t"
        mode = "spatial" if which == "S" else "temporal"
        x = x + _attention(ln(x, "norm1_" + s), P, p + "attn_" + s, mode, F, cfg)
        x = x + _mlp(ln(x, "norm2_" + s), P, p + "mlp_" + s)
    return x


def forward(P: "dict[str, np.ndarray]", x: np.ndarray, cfg: EncoderConfig,
            dtype=np.float32, return_intermediates: bool = False):
    """DSTformer.forward (DSTformer.py:329-358).  Returns (out, rep):
    out = forward(x), rep = get_representation(x) (:360-361)."""
    P = {k: v.astype(dtype) for k, v in P.items()}
    x = x.astype(dtype)
    B, F, J, _ = x.shape
    C = cfg.dim_feat
    inter = {}
    h = x.reshape(B * F, J, cfg.dim_in)                                     # :331
    h = _linear(h, P["joints_embed.weight"], P["joints_embed.bias"])        # :333
    h = h + P["pos_embed"]                                                  # :334
    h = h.reshape(B, F, J, C) + P["temp_embed"][:, :F]                      # :336
    h = h.reshape(B * F, J, C)                                              # :337
    inter["embed"] = h
    for i in range(cfg.depth):                                              # :340
        x_st = _block(h, P, f"blocks_st.{i}.", "ST", F, cfg)                # :341
        x_ts = _block(h, P, f"blocks_ts.{i}.", "TS", F, cfg)                # :342
        a = _linear(np.concatenate([x_st, x_ts], axis=-1),
                    P[f"ts_attn.{i}.weight"], P[f"ts_attn.{i}.bias"])       # :345-347
        a = _softmax(a)                                                     # :348
        h = x_st * a[..., 0:1] + x_ts * a[..., 1:2]                         # :349
        inter[f"depth{i}"] = h
    h = _layer_norm(h, P["norm.weight"], P["norm.bias"], cfg.eps)           # :352
    h = h.reshape(B, F, J, C)                                               # :353
    rep = np.tanh(_linear(h, P["pre_logits.fc.weight"], P["pre_logits.fc.bias"]))  # :354
    out = _linear(rep, P["head.weight"], P["head.bias"])                    # :357
    if return_intermediates:
        return out, rep, inter
    return out, rep


# --------------------------------------------------------------------------- #
# algorithmic work (SURVEY.md section 0 / BASELINE.md section 4)
# --------------------------------------------------------------------------- #
def flops_per_sequence(cfg: EncoderConfig, T: int) -> float:
    C, hid, J = cfg.dim_feat, cfg.hidden, cfg.num_joints
    lin = 8 * C * C
    mlp = 4 * C * hid
    s = 4 * J * C
    t = 4 * T * C
    fuse = 8 * C
    tok = cfg.depth * (4 * lin + 4 * mlp + 2 * s + 2 * t + fuse) \
        + 2 * cfg.dim_in * C + 2 * C * cfg.dim_rep + 2 * cfg.dim_rep * cfg.dim_out
    return float(tok) * T * J


def mpjpe(pred: np.ndarray, gt: np.ndarray) -> float:
    """lib/model/loss.py:8-14 -- mean per-joint position error."""
    return float(np.mean(np.linalg.norm(pred - gt, axis=-1)))
