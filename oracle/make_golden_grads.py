"""Generate tests/golden/grads_*.npz from the REAL reference  --  run in the build container only.

    python oracle/make_golden_grads.py          # needs /root/reference (read-only mount)

Gradient fixtures for SURVEY.md section 8 row a15: the unmodified reference module
(`/root/reference/lib/model/DSTformer.py`, built by its own factory `lib/utils/learning.py:79-85`) in float64, the
deterministic perturbed parameters / clips of `oracle.dstformer_oracle`, loss = sum(out * w) with a seeded w, and
`torch.autograd` -- i.e. exactly what `train.py:205 loss_total.backward()` differentiates.  Per parameter tensor the
fixture keeps ||g||_2, sum(g) and up to 512 evenly spaced entries (42.5 M float64 gradients would not be a "small
fixture"); the gradient w.r.t. the input clip is kept whole.  The same run pins the oracle's differentiable restatement
(`oracle/dstformer_torch_autograd.py`): its float64 gradients are compared with the reference's and the max deviation is
written to tests/golden/MANIFEST_GRADS.json.  No reference source is copied: only numbers the reference produced.
"""
from __future__ import annotations

import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import dstformer_oracle as O                     # noqa: E402
from oracle.dstformer_torch_autograd import recompute_forward  # noqa: E402
from lib.utils.learning import load_backbone                 # noqa: E402  (the reference factory)

GOLD = os.path.join(ROOT, "tests", "golden")
NSAMPLE = 512

# (name, cfg, B, F, param_seed, input_seed, w_seed, return_rep)
CASES = [
    ("grads_base_b2_f27", O.BASE, 2, 27, 11, 1, 101, False),
    ("grads_lite_b2_f27", O.LITE, 2, 27, 21, 5, 102, False),
    ("grads_base_b1_f243", O.BASE, 1, 243, 11, 2, 103, False),
    ("grads_lite_b2_f40_rep", O.LITE, 2, 40, 22, 7, 104, True),      # get_representation path (train_action / train_mesh)
]


def sample_idx(n: int) -> np.ndarray:
    return O.fixture_sample_idx(n, NSAMPLE)


def out_weight(shape, seed: int) -> np.ndarray:
    return O.fixture_out_weight(shape, seed)


def ordered_names(cfg) -> "list[str]":
    """Parameter names in `DSTformer._ordered_params()` / `mb_param_info` order (== O.param_shapes order)."""
    return list(O.param_shapes(cfg).keys())


def main():
    torch.set_num_threads(os.cpu_count())
    manifest = {"reference": "Walter0807/MotionBERT @ /root/reference (lib/model/DSTformer.py), float64 autograd",
                "torch": torch.__version__, "cases": {}}
    for name, cfg, B, F, ps, xs, ws, return_rep in CASES:
        P = O.make_params(cfg, ps)
        x = O.make_input(B, F, cfg.num_joints, xs)
        args = SimpleNamespace(backbone="DSTformer", dim_feat=cfg.dim_feat, dim_rep=cfg.dim_rep, depth=cfg.depth,
                               num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio, maxlen=cfg.maxlen,
                               num_joints=cfg.num_joints)
        m = load_backbone(args).double().train()        # train(): all drop rates are 0 in the factory's build
        m.load_state_dict({k: torch.from_numpy(v).double() for k, v in P.items()}, strict=True)
        xt = torch.from_numpy(x).double().requires_grad_(True)
        y = m.get_representation(xt) if return_rep else m(xt)
        w = torch.from_numpy(out_weight(tuple(y.shape), ws))
        (y * w).sum().backward()
        named = dict(m.named_parameters())
        names = ordered_names(cfg)
        store = {"B": B, "F": F, "param_seed": ps, "input_seed": xs, "w_seed": ws, "return_rep": int(return_rep),
                 "dim_feat": cfg.dim_feat, "mlp_ratio": cfg.mlp_ratio,
                 "loss": np.float64(float((y * w).sum())), "dx": xt.grad.numpy()}
        ref_grads = []
        for i, n in enumerate(names):
            g = named[n].grad
            g = torch.zeros_like(named[n]) if g is None else g      # head.* on the representation path
            ref_grads.append(g)
            flat = g.reshape(-1).numpy()
            idx = sample_idx(flat.size)
            store[f"idx_{i}"] = idx
            store[f"val_{i}"] = flat[idx]
            store[f"norm_{i}"] = np.float64(np.linalg.norm(flat))
            store[f"sum_{i}"] = np.float64(flat.sum())
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **store)
        # pin the oracle's differentiable restatement against the reference's own gradients
        mod = SimpleNamespace(dim_feat=cfg.dim_feat, num_heads=cfg.num_heads, qk_scale=None, depth=cfg.depth,
                              att_fuse=True, eps=cfg.eps)
        pl = [torch.from_numpy(P[n]).double().requires_grad_(True) for n in names]
        xr = torch.from_numpy(x).double().requires_grad_(True)
        yr = recompute_forward(mod, xr, return_rep, None, pl)
        gr = torch.autograd.grad((yr * w).sum(), [xr] + pl, allow_unused=True)
        dev = 0.0
        for g_ref, g_or in zip([xt.grad] + ref_grads, gr):
            g_or = torch.zeros_like(g_ref) if g_or is None else g_or
            den = float(g_ref.norm())
            if den > 0:
                dev = max(dev, float((g_or - g_ref).norm()) / den)
        manifest["cases"][name] = {"B": B, "F": F, "return_rep": return_rep, "n_tensors": len(names),
                                   "loss": float((y * w).sum()),
                                   "forward_restatement_vs_ref_maxabs": float((yr - y).abs().max()),
                                   "oracle_autograd_vs_ref_max_rel_l2": dev}
        print(name, json.dumps(manifest["cases"][name]), flush=True)
    with open(os.path.join(GOLD, "MANIFEST_GRADS.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
