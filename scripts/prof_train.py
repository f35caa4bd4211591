"""One profiled training step for ncu (--profile-from-start off): bf16 forward_train + fused loss + native backward."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, synthetic_clips  # noqa: E402
from motionbert_b200.loss import pretrain_loss_3d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="base")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=243)
ap.add_argument("--math", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
m = build_model(a.model, dev, a.math).train()
x = synthetic_clips(a.batch, a.frames, 1).to(dev)
gt = synthetic_clips(a.batch, a.frames, 2).to(dev)


def step():
    m.zero_grad(set_to_none=True)
    total, _ = pretrain_loss_3d(m(x), gt, 0.5, 20.0)
    total.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
