"""Small-batch latency: eager (Python -> ctypes -> 108 launches) vs CUDA-graph replay of the same forward."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, synthetic_clips  # noqa: E402

dev = torch.device("cuda:0")
for model, B, T in (("base", 1, 243), ("base", 2, 243), ("lite", 1, 27), ("lite", 1, 243), ("base", 8, 16)):
    m = build_model(model, dev, "bf16x3")
    x = synthetic_clips(B, T, 1).to(dev)
    with torch.no_grad():
        for _ in range(5):
            m(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            m(x)
        torch.cuda.synchronize()
        eager = (time.perf_counter() - t0) / 50
        run = m.make_graphed(B, T)
        for _ in range(5):
            run(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            run(x)
        torch.cuda.synchronize()
        graphed = (time.perf_counter() - t0) / 50
    print(f"{model} B={B} T={T}: eager {eager * 1e3:.3f} ms  graph {graphed * 1e3:.3f} ms")
