"""CPU oracle for `Augmenter2D` (lib/data/augmentation.py:29-74)  --  TEST INFRASTRUCTURE ONLY.

numpy restatement of `add_noise` / `add_mask` given the random draws the reference makes (in its order: sel, gauss,
unif, jitter [augmentation.py:43-47], shift [:24], then mask, mask_T [:71-72]).  Pinned against the real module by
`oracle/make_golden_augment.py` -> tests/golden/augment2d.npz.  Never imported by the product package.
"""
from __future__ import annotations

import numpy as np


def add_noise(x, sel, gauss, unif, jitter, shift, mean, std, weight, a, b, m, s, uniform_range=0.06, noise_std=0.002):
    """x (B,F,J,>=2) -> (B,F,J,3).  augmentation.py:29-65."""
    x = np.asarray(x, np.float64)[..., :2]
    B, F, J, _ = x.shape
    K = sel.shape[1]
    sel = np.asarray(sel, np.float64).reshape(B, K, J, 1)
    g = np.asarray(gauss, np.float64) * np.asarray(std, np.float64) + np.asarray(mean, np.float64)      # :44
    u = (np.asarray(unif, np.float64) - 0.5) * uniform_range                                              # :45
    w = np.asarray(weight, np.float64).reshape(1, 1, J, 1)
    delta = g * (sel < w) + u * (sel >= w)                                                                # :57
    # F.interpolate(..., [F, J, 2], mode='trilinear', align_corners=True): only the frame axis is resampled  (:58)
    src = np.arange(F, dtype=np.float64) * ((K - 1) / (F - 1) if F > 1 else 0.0)
    k0 = np.floor(src).astype(np.int64)
    k1 = np.minimum(k0 + 1, K - 1)
    l1 = (src - k0).reshape(1, F, 1, 1)
    expand = delta[:, k0] * (1.0 - l1) + delta[:, k1] * l1
    final = expand + np.asarray(jitter, np.float64)[None] * noise_std                                    # :47,59
    xy = x + final                                                                                        # :60
    dis = np.sqrt((final ** 2).sum(-1))                                                                   # :61-64
    conf = a / (dis + a) + b * dis + (np.asarray(shift, np.float64) * s + m)                              # :22-27
    return np.concatenate([xy, np.clip(conf, 0.0, 1.0)[..., None]], axis=-1)                              # :65-66


def add_mask(x, mask_u, maskT_u, mask_ratio, mask_T_ratio):
    """augmentation.py:67-74."""
    x = np.asarray(x, np.float64)
    B, F, J, _ = x.shape
    mk = (np.asarray(mask_u).reshape(B, F, J, 1) > mask_ratio).astype(np.float64)
    mt = (np.asarray(maskT_u).reshape(1, F, 1, 1) > mask_T_ratio).astype(np.float64)
    return x * mk * mt
