"""Tiny forward of both model sizes for compute-sanitizer (memcheck / synccheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import build_module, rel_token_err  # noqa: E402
from oracle import dstformer_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
for cfg, B, F in ((O.LITE, 2, 9), (O.BASE, 1, 131)):
    P = O.make_params(cfg, 3)
    x = O.make_input(B, F, 17, 4)
    m = build_module(cfg, P, dev)
    with torch.no_grad():
        out = m(torch.from_numpy(x).to(dev)).cpu().numpy()
    o_ref, _ = O.forward(P, x, cfg, np.float32)
    print(cfg.dim_feat, B, F, "rel", rel_token_err(out, o_ref))
print("sanitize forward done")
