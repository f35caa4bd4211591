"""One profiled forward for ncu (run with --profile-from-start off): warm up, then cudaProfilerStart/Stop
around a single forward of the product path."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, synthetic_clips  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="base")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=243)
ap.add_argument("--math", default="f16c")
ap.add_argument("--kernel-flags", type=lambda v: int(v, 0), default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
m = build_model(a.model, dev, a.math)
m._kernel_flags = a.kernel_flags
x = synthetic_clips(a.batch, a.frames, 1).to(dev)
with torch.no_grad():
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
