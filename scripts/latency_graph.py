"""Small-clip latency of DSTformer.forward (infer_wild-style calls): eager launches vs the automatic CUDA-graph replay
(motionbert_b200.DSTformer._auto_graph, default for B*F*J <= 16384 tokens).  Wall clock per call incl. the Python host path,
result synchronised every call (latency, not throughput)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, synthetic_clips  # noqa: E402

dev = torch.device("cuda:0")


def timed(m, x, n=200):
    with torch.no_grad():
        for _ in range(10):
            m(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            m(x)
            torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3


for model, B, T in (("lite", 1, 27), ("lite", 1, 243), ("base", 1, 27), ("base", 1, 243), ("base", 2, 243), ("base", 8, 16)):
    m = build_model(model, dev, "f16c")
    x = synthetic_clips(B, T, 1).to(dev)
    m.auto_graph_max_tokens = 0
    eager = timed(m, x)
    m.auto_graph_max_tokens = 16384
    graphed = timed(m, x)
    print(json.dumps({"model": model, "B": B, "T": T, "tokens": B * T * 17, "eager_ms": round(eager, 4), "auto_graph_ms": round(graphed, 4)}))
