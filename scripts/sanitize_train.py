"""Tiny training steps (forward_train + fused loss + native backward) for compute-sanitizer memcheck."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_backward import _module  # noqa: E402
from motionbert_b200.loss import pretrain_loss_3d  # noqa: E402
from oracle import dstformer_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
for dim, heads, mlp, B, F in ((256, 8, 2, 2, 9), (512, 8, 2, 1, 131), (256, 4, 4, 3, 33)):
    m = _module(dev, dim, 1, heads, mlp, seed=1)
    x = torch.from_numpy(O.make_input(B, F, 17, 4)).to(dev)
    gt = torch.from_numpy(O.make_input(B, F, 17, 5)).to(dev)
    total, _ = pretrain_loss_3d(m(x), gt, 0.5, 20.0)
    total.backward()
    torch.cuda.synchronize()
    print(dim, heads, B, F, "loss", float(total.detach()), "grad norm", float(sum(p.grad.norm() ** 2 for p in m.parameters()) ** 0.5))
print("sanitize train done")
