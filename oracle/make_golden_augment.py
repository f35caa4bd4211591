"""Generate tests/golden/augment2d.npz from the REAL `Augmenter2D`  --  run in the build container only.

    python oracle/make_golden_augment.py          # needs /root/reference (read-only mount)

Runs the unmodified `lib/data/augmentation.py::Augmenter2D` (with the reference's own params/*.pkl/.pth and the shipped
mask ratios, configs/pretrain/MB_pretrain.yaml:49-50) on seeded clips, captures the random draws by replaying the same
seeded torch calls in the reference's order, and stores inputs, draws, noise-model constants and the reference outputs.
`lib/utils/tools.py` imports `easydict`, which is not installed here: a stub module stands in for the import only.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
if "easydict" not in sys.modules:                      # import-only stub (get_config is never called)
    stub = types.ModuleType("easydict")
    stub.EasyDict = dict
    sys.modules["easydict"] = stub

from lib.data.augmentation import Augmenter2D          # noqa: E402  (the real module)
from oracle import augment_oracle as AO               # noqa: E402
from oracle import dstformer_oracle as O              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    args = SimpleNamespace(d2c_params_path="/root/reference/params/d2c_params.pkl",
                           noise_path="/root/reference/params/synthetic_noise.pth", mask_ratio=0.05, mask_T_ratio=0.1)
    aug = Augmenter2D(args)
    store = {"mean": aug.noise["mean"].numpy(), "std": aug.noise["std"].numpy(), "weight": aug.noise["weight"].numpy(),
             "a": np.float64(aug.d2c_params["a"]), "b": np.float64(aug.d2c_params["b"]),
             "m": np.float64(aug.d2c_params["m"]), "s": np.float64(aug.d2c_params["s"]),
             "mask_ratio": np.float64(args.mask_ratio), "mask_T_ratio": np.float64(args.mask_T_ratio)}
    cases = [("a", 3, 243, 5), ("b", 2, 27, 6), ("c", 4, 1, 7), ("d", 2, 81, 8)]
    store["cases"] = np.array([c[0] for c in cases])
    for name, B, F, seed in cases:
        x = torch.from_numpy(O.make_input(B, F, 17, seed))
        K, J = aug.num_Kframes, 17
        torch.manual_seed(seed)                                             # replay: the reference's draws, in its order
        sel = torch.rand((B, K, J, 1)); gauss = torch.randn(B, K, J, 2); unif = torch.rand((B, K, J, 2))
        jitter = torch.randn(F, J, 2); shift = torch.randn(B, F, J)
        mask_u = torch.rand(B, F, J, 1); maskT_u = torch.rand(1, F, 1, 1)
        torch.manual_seed(seed)
        both = aug.augment2D(x, mask=True, noise=True)
        torch.manual_seed(seed)
        noise_only = aug.augment2D(x, mask=False, noise=True)
        # mask-only: the mask draws come first in that call
        torch.manual_seed(seed + 100)
        mu2 = torch.rand(B, F, J, 1); mt2 = torch.rand(1, F, 1, 1)
        torch.manual_seed(seed + 100)
        mask_only = aug.augment2D(x, mask=True, noise=False)
        c = dict(mean=store["mean"], std=store["std"], weight=store["weight"], a=float(store["a"]), b=float(store["b"]),
                 m=float(store["m"]), s=float(store["s"]))
        o_noise = AO.add_noise(x.numpy(), sel.numpy(), gauss.numpy(), unif.numpy(), jitter.numpy(), shift.numpy(), **c)
        o_both = AO.add_mask(o_noise, mask_u.numpy(), maskT_u.numpy(), args.mask_ratio, args.mask_T_ratio)
        o_mask = AO.add_mask(x.numpy(), mu2.numpy(), mt2.numpy(), args.mask_ratio, args.mask_T_ratio)
        dev = max(np.abs(o_noise - noise_only.numpy()).max(), np.abs(o_both - both.numpy()).max(),
                  np.abs(o_mask - mask_only.numpy()).max())
        print(f"case {name}: B={B} F={F} oracle vs real Augmenter2D max abs {dev:.2e}")
        assert dev < 5e-6, "draw replay does not reproduce the reference"
        for k, v in dict(x=x, sel=sel, gauss=gauss, unif=unif, jitter=jitter, shift=shift, mask_u=mask_u, maskT_u=maskT_u,
                         mask_u2=mu2, maskT_u2=mt2, out_both=both, out_noise=noise_only, out_mask=mask_only).items():
            store[f"{name}_{k}"] = v.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "augment2d.npz"), **store)


if __name__ == "__main__":
    main()
