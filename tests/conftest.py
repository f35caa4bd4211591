"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only build box; `-m gpu` needs one B200."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def golden_names():
    with open(os.path.join(GOLD, "MANIFEST.json")) as f:
        return sorted(json.load(f)["cases"].keys())


def manifest():
    with open(os.path.join(GOLD, "MANIFEST.json")) as f:
        return json.load(f)


def load_case(name):
    """Returns (cfg, params(dict of np.float32), x, golden npz) for a committed fixture."""
    from oracle import dstformer_oracle as O
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = O.EncoderConfig(dim_feat=int(g["dim_feat"]), mlp_ratio=float(g["mlp_ratio"]))
    P = O.make_params(cfg, int(g["param_seed"]))
    x = O.make_input(int(g["B"]), int(g["F"]), cfg.num_joints, int(g["input_seed"]))
    return cfg, P, x, g


def build_module(cfg, P=None, device=None):
    """The product class constructed the way lib/utils/learning.py:83-85 does it."""
    from functools import partial

    import torch
    import torch.nn as nn

    from motionbert_b200 import DSTformer
    m = DSTformer(dim_in=3, dim_out=3, dim_feat=cfg.dim_feat, dim_rep=cfg.dim_rep, depth=cfg.depth,
                  num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                  maxlen=cfg.maxlen, num_joints=cfg.num_joints)
    if P is not None:
        m.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()}, strict=True)
    if device is not None:
        m = m.to(device)
    return m.eval()


def rel_token_err(a, b):
    """per-token relative L2 error (the north-star's 1e-3 criterion), returns (mean, max)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, a.shape[-1])
    b = np.asarray(b, dtype=np.float64).reshape(-1, b.shape[-1])
    e = np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-12)
    return float(e.mean()), float(e.max())


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
