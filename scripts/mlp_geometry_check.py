"""One-off GPU check (not collected by pytest): fused vs two-GEMM MLP, bit for bit, on tile geometries no shipped config uses."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_module
from motionbert_b200 import _lib
from oracle import dstformer_oracle as O
dev = torch.device("cuda:0")
ok = True
for dim, ratio, B, F in ((512, 4, 3, 50), (256, 1, 3, 50), (256, 2, 2, 33), (512, 1, 2, 40), (512, 4, 8, 243)):
    cfg = O.EncoderConfig(dim_feat=dim, mlp_ratio=ratio, depth=2)
    P = O.make_params(cfg, 5)
    x = torch.from_numpy(O.make_input(B, F, 17, 9)).to(dev)
    m = build_module(cfg, P, dev)
    with torch.no_grad():
        m._kernel_flags = _lib.MB_FLAG_MLP_SPLIT; r_s = m.get_representation(x).clone()
        m._kernel_flags = 0; r_f = m.get_representation(x).clone()
        m._kernel_flags = _lib.MB_FLAG_MLP_NO_RING; r_n = m.get_representation(x).clone()
    o_ref, r_ref = O.forward(P, x.cpu().numpy()[:1], cfg, np.float64) if B * F < 400 else (None, None)
    e = torch.equal(r_f, r_s) and torch.equal(r_n, r_s) and bool(torch.isfinite(r_f).all())
    rel = None
    if r_ref is not None:
        d = r_f[:1].cpu().numpy().astype(np.float64) - r_ref
        rel = float((np.linalg.norm(d, axis=-1) / np.linalg.norm(r_ref, axis=-1)).max())
    print(f"C={dim} hidden={int(dim*ratio)} NT1={int(dim*ratio)//256} NT2={dim//256} B={B} F={F}: fused==split {e}  max per-token rel vs fp64 oracle {rel}")
    ok = ok and e
print("ALL OK" if ok else "MISMATCH")
