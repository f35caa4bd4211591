"""Training-step timing (forward_train + backward) of the drop-in module on one GPU: native backward vs the
torch-op recompute fallback.  Not the headline bench (BASELINE's metric is the forward); evidence for row a15."""
import argparse
import json
import os
import sys
from functools import partial

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motionbert_b200 import DSTformer  # noqa: E402


def run(B, F, steps, warmup, torch_bwd, lite):
    if torch_bwd:
        os.environ["MB_TORCH_BACKWARD"] = "1"
    else:
        os.environ.pop("MB_TORCH_BACKWARD", None)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    kw = dict(dim_feat=256, depth=5, num_heads=8, mlp_ratio=4) if lite else dict(dim_feat=512, depth=5, num_heads=8, mlp_ratio=2)
    m = DSTformer(dim_in=3, dim_out=3, dim_rep=512, maxlen=243, norm_layer=partial(nn.LayerNorm, eps=1e-6), **kw).to(dev).train()
    x = torch.randn(B, F, 17, 3, device=dev)
    w = torch.randn(B, F, 17, 3, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd = bwd = 0.0
    for it in range(warmup + steps):
        m.zero_grad(set_to_none=True)
        ev[0].record()
        out = m(x)
        loss = (out * w).sum()
        ev[1].record()
        loss.backward()
        ev[2].record()
        torch.cuda.synchronize()
        if it >= warmup:
            fwd += ev[0].elapsed_time(ev[1])
            bwd += ev[1].elapsed_time(ev[2])
    return fwd / steps, bwd / steps


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=243)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--lite", action="store_true")
    ap.add_argument("--skip-torch", action="store_true")
    a = ap.parse_args()
    f, b = run(a.batch, a.frames, a.steps, a.warmup, False, a.lite)
    res = {"config": {"model": "lite" if a.lite else "base", "B": a.batch, "F": a.frames}, "native": {"fwd_ms": f, "bwd_ms": b,
           "seq_per_s": a.batch / ((f + b) / 1e3)}}
    if not a.skip_torch:
        f2, b2 = run(a.batch, a.frames, a.steps, a.warmup, True, a.lite)
        res["torch_recompute_backward"] = {"fwd_ms": f2, "bwd_ms": b2, "seq_per_s": a.batch / ((f2 + b2) / 1e3)}
    print(json.dumps(res))
