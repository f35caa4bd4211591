"""Generate tests/golden/*.npz from the REAL reference  --  run in the build container only.

    python oracle/make_golden.py          # needs /root/reference (read-only mount)

Imports the unmodified reference module (`/root/reference/lib/model/DSTformer.py`)
through its own factory (`lib/utils/learning.py:79-85`), loads the deterministic
perturbed parameters of `oracle.dstformer_oracle.make_params`, runs it on the
deterministic synthetic clips of `make_input`, and stores inputs-by-seed +
outputs as small fixtures.  It also pins the oracle: the numpy and torch-CPU
restatements are run on the same inputs and their max deviation from the
reference is written to tests/golden/MANIFEST.json (the `-m "not gpu"` tests
re-check the oracle against the stored reference outputs on every box).

No reference source is copied: only numbers the reference produced.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import dstformer_oracle as O            # noqa: E402
from oracle import dstformer_torch_cpu as OT        # noqa: E402
from lib.utils.learning import load_backbone        # noqa: E402  (the reference factory)

GOLD = os.path.join(ROOT, "tests", "golden")

# (name, cfg, B, F, param_seed, input_seed)
CASES = [
    ("base_b2_f27", O.BASE, 2, 27, 11, 1),
    ("base_b1_f243", O.BASE, 1, 243, 11, 2),
    ("base_b3_f16", O.BASE, 3, 16, 12, 3),
    ("base_b1_f1", O.BASE, 1, 1, 11, 4),
    ("base_b2_f130", O.BASE, 2, 130, 12, 9),
    ("lite_b2_f27", O.LITE, 2, 27, 21, 5),
    ("lite_b1_f243", O.LITE, 1, 243, 21, 6),
    ("lite_b2_f81", O.LITE, 2, 81, 22, 7),
    ("lite_b5_f30", O.LITE, 5, 30, 22, 8),
]


def ref_model(cfg: O.EncoderConfig):
    args = SimpleNamespace(backbone="DSTformer", dim_feat=cfg.dim_feat, dim_rep=cfg.dim_rep,
                           depth=cfg.depth, num_heads=cfg.num_heads, mlp_ratio=cfg.mlp_ratio,
                           maxlen=cfg.maxlen, num_joints=cfg.num_joints)
    return load_backbone(args)


def token_sample(M: int, n: int = 64) -> np.ndarray:
    return np.unique(np.linspace(0, M - 1, num=min(n, M)).round().astype(np.int64))


def init_checksums(cfg):
    """Known-answer values of the reference's own init."""
    torch.manual_seed(0)
    m = ref_model(cfg).eval()
    sd = m.state_dict()
    tot = sum(float(v.double().sum()) for v in sd.values())
    tot2 = sum(float((v.double() ** 2).sum()) for v in sd.values())
    sha = hashlib.sha256(b"".join(v.contiguous().numpy().tobytes() for v in sd.values())).hexdigest()[:16]
    x = torch.rand(2, 27, 17, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1
    with torch.no_grad():
        out = m(x)
        rep = m.get_representation(x)
    return {"n_tensors": len(sd), "sum": tot, "sumsq": tot2, "sha16": sha,
            "out_sum": float(out.double().sum()), "out_000": [float(v) for v in out[0, 0, 0]],
            "rep_sum": float(rep.double().sum()),
            "rep_last3": [float(v) for v in rep[1, 26, 16, :3]]}


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    manifest = {"reference": "Walter0807/MotionBERT @ /root/reference (lib/model/DSTformer.py)",
                "torch": torch.__version__, "numpy": np.__version__, "cases": {}, "init_kat": {}}
    for name, cfg, B, F, ps, xs in CASES:
        P = O.make_params(cfg, ps)
        x = O.make_input(B, F, cfg.num_joints, xs)
        m = ref_model(cfg).eval()
        m.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()}, strict=True)
        with torch.no_grad():
            out = m(torch.from_numpy(x)).numpy()
            rep = m.get_representation(torch.from_numpy(x)).numpy()
            m64 = ref_model(cfg).double().eval()
            m64.load_state_dict({k: torch.from_numpy(v).double() for k, v in P.items()}, strict=True)
            out64 = m64(torch.from_numpy(x).double()).numpy()
            rep64 = m64.get_representation(torch.from_numpy(x).double()).numpy()
        # pin the oracle restatements against the real reference
        o_np, r_np = O.forward(P, x, cfg, np.float32)
        o_np64, r_np64 = O.forward(P, x, cfg, np.float64)
        o_t, r_t = OT.forward({k: torch.from_numpy(v) for k, v in P.items()}, torch.from_numpy(x),
                              cfg.depth, cfg.num_heads, cfg.eps)
        M = B * F * cfg.num_joints
        idx = token_sample(M)
        np.savez_compressed(
            os.path.join(GOLD, name + ".npz"),
            B=B, F=F, param_seed=ps, input_seed=xs, dim_feat=cfg.dim_feat, mlp_ratio=cfg.mlp_ratio,
            out=out.astype(np.float32), out64=out64.astype(np.float64),
            rep_idx=idx, rep_rows=rep.reshape(M, -1)[idx].astype(np.float32),
            rep_rows64=rep64.reshape(M, -1)[idx].astype(np.float64),
            rep_sum=np.float64(rep.astype(np.float64).sum()),
            rep_absmean=np.float64(np.abs(rep).mean()),
        )
        manifest["cases"][name] = {
            "B": B, "F": F, "param_seed": ps, "input_seed": xs,
            "dim_feat": cfg.dim_feat, "mlp_ratio": cfg.mlp_ratio,
            "out_absmean": float(np.abs(out).mean()), "rep_absmean": float(np.abs(rep).mean()),
            "ref32_vs_ref64_out_maxabs": float(np.abs(out - out64).max()),
            "oracle_np32_vs_ref_out_maxabs": float(np.abs(o_np - out).max()),
            "oracle_np32_vs_ref_rep_maxabs": float(np.abs(r_np - rep).max()),
            "oracle_np64_vs_ref64_out_maxabs": float(np.abs(o_np64 - out64).max()),
            "oracle_np64_vs_ref64_rep_maxabs": float(np.abs(r_np64 - rep64).max()),
            "oracle_torch_vs_ref_out_maxabs": float(np.abs(o_t.numpy() - out).max()),
            "oracle_torch_vs_ref_rep_maxabs": float(np.abs(r_t.numpy() - rep).max()),
        }
        print(name, json.dumps(manifest["cases"][name]))
    manifest["init_kat"]["base"] = init_checksums(O.BASE)
    manifest["init_kat"]["lite"] = init_checksums(O.LITE)
    print(json.dumps(manifest["init_kat"], indent=1))
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
