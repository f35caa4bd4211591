"""Augmenter2D (lib/data/augmentation.py:29-74): the oracle against outputs of the REAL
module (tests/golden/augment2d.npz, made by oracle/make_golden_augment.py), and the one-kernel GPU version against both."""
import os
import pickle
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import augment_oracle as AO


def _fx():
    return np.load(os.path.join(GOLD, "augment2d.npz"))


def _consts(g):
    return dict(mean=g["mean"], std=g["std"], weight=g["weight"], a=float(g["a"]), b=float(g["b"]), m=float(g["m"]), s=float(g["s"]))


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_oracle_matches_the_real_augmenter(case):
    g = _fx()
    c = _consts(g)
    k = lambda n: g[f"{case}_{n}"]   # noqa: E731
    o_noise = AO.add_noise(k("x"), k("sel"), k("gauss"), k("unif"), k("jitter"), k("shift"), **c)
    np.testing.assert_allclose(o_noise, k("out_noise"), atol=5e-6, rtol=0)
    o_both = AO.add_mask(o_noise, k("mask_u"), k("maskT_u"), float(g["mask_ratio"]), float(g["mask_T_ratio"]))
    np.testing.assert_allclose(o_both, k("out_both"), atol=5e-6, rtol=0)
    o_mask = AO.add_mask(k("x"), k("mask_u2"), k("maskT_u2"), float(g["mask_ratio"]), float(g["mask_T_ratio"]))
    np.testing.assert_allclose(o_mask, k("out_mask"), atol=0, rtol=0)
    assert (o_both == 0).mean() > 0.05                      # the masks do remove joints / frames


def _augmenter(tmp_path, g):
    from motionbert_b200.augment import Augmenter2D
    with open(tmp_path / "d2c.pkl", "wb") as f:
        pickle.dump({"a": float(g["a"]), "b": float(g["b"]), "m": np.float32(g["m"]), "s": np.float32(g["s"])}, f)
    torch.save({"mean": torch.from_numpy(g["mean"]), "std": torch.from_numpy(g["std"]), "weight": torch.from_numpy(g["weight"])},
               tmp_path / "noise.pth")
    return Augmenter2D(SimpleNamespace(d2c_params_path=str(tmp_path / "d2c.pkl"), noise_path=str(tmp_path / "noise.pth"),
                                       mask_ratio=float(g["mask_ratio"]), mask_T_ratio=float(g["mask_T_ratio"])))


def test_cpu_tensor_fails_loudly(tmp_path):
    aug = _augmenter(tmp_path, _fx())
    with pytest.raises(RuntimeError, match="CUDA"):
        aug.add_mask(torch.zeros(1, 2, 17, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_gpu_kernel_matches_the_real_augmenter(cuda_device, tmp_path, case):
    g = _fx()
    aug = _augmenter(tmp_path, g)
    t = lambda n: torch.from_numpy(g[f"{case}_{n}"]).to(cuda_device)   # noqa: E731
    x = t("x")
    draws = (t("sel"), t("gauss"), t("unif"), t("jitter"), t("shift"))
    out_noise = aug._launch(x, True, False, draws=draws)
    out_both = aug._launch(x, True, True, draws=draws, mask_draws=(t("mask_u"), t("maskT_u")))
    out_mask = aug._launch(x, False, True, mask_draws=(t("mask_u2"), t("maskT_u2")))
    np.testing.assert_allclose(out_noise.cpu().numpy(), g[f"{case}_out_noise"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(out_both.cpu().numpy(), g[f"{case}_out_both"], atol=5e-6, rtol=0)
    np.testing.assert_allclose(out_mask.cpu().numpy(), g[f"{case}_out_mask"], atol=0, rtol=0)


@pytest.mark.gpu
def test_same_seed_reproduces_the_reference_noise(cuda_device, tmp_path):
    """add_noise draws on the CPU generator in the reference's order: the reference's seed gives the reference's clip."""
    g = _fx()
    aug = _augmenter(tmp_path, g)
    x = torch.from_numpy(g["b_x"]).to(cuda_device)
    torch.manual_seed(6)                                    # case b was generated with seed 6
    out = aug.augment2D(x, mask=False, noise=True)
    np.testing.assert_allclose(out.cpu().numpy(), g["b_out_noise"], atol=5e-6, rtol=0)
    assert aug.augment2D(x) is x                             # neither stage: the input is returned untouched
