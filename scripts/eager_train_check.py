"""One-off check of bench.gpu_eager_train_baseline on the GPU box (config 3's torch-eager comparator)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
m = bench.build_model("base", dev, "f16c")
print(json.dumps(bench.gpu_eager_train_baseline(m, "base", 243, dev)))
print("max memory GB", torch.cuda.max_memory_allocated() / 1e9)
