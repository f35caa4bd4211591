"""`Augmenter2D` of the pretraining loop on the GPU in one kernel.

Mirror of `lib/data/augmentation.py:10-81` (class name, constructor argument `args` with `d2c_params_path`, `noise_path`,
`mask_ratio`, `mask_T_ratio`; methods `add_noise`, `add_mask`, `augment2D`), called by `train.py:162-172` right before
the encoder.  The random numbers are drawn exactly like the reference draws them -- same calls, same order, same shapes,
on the CPU generator for `add_noise` and on the input's device for `add_mask` -- so a seed reproduces the reference's
augmentation; everything after the draws (Gaussian / uniform selection per key frame, linear interpolation over the
frames, jitter, confidence re-synthesis, clipping, masking, concatenation: ~25 elementwise kernels and a trilinear
interpolate in the reference) is ONE launch of `mb_augment2d`.  CUDA tensors only; no CPU fallback.
"""
from __future__ import annotations

import ctypes
import pickle

import torch

from . import _lib


class Augmenter2D(object):
    def __init__(self, args):
        with open(args.d2c_params_path, "rb") as f:                    # lib/utils/tools.py read_pkl
            self.d2c_params = pickle.load(f)
        self.noise = torch.load(args.noise_path)
        self.mask_ratio = args.mask_ratio
        self.mask_T_ratio = args.mask_T_ratio
        self.num_Kframes = 27
        self.noise_std = 0.002
        self._dev = {}

    def _consts(self, device):
        c = self._dev.get(device)
        if c is None:
            c = tuple(self.noise[k].float().contiguous().to(device) for k in ("mean", "std", "weight"))
            self._dev[device] = c
        return c

    def _launch(self, x, noise, mask, draws=None, mask_draws=None):
        if not x.is_cuda:
            raise RuntimeError("motionbert_b200.augment runs on sm_100a CUDA devices only (no CPU fallback)")
        B, F, J, cin = x.shape
        xc = x.detach().float().contiguous()
        out = torch.empty(B, F, J, 3, dtype=torch.float32, device=x.device)
        mean, std, weight = self._consts(x.device)
        d = self.d2c_params
        ur = float(self.noise["uniform_range"]) if "uniform_range" in self.noise.keys() else 0.06
        sel = gauss = unif = jitter = shift = mk = mt = None
        if noise:
            sel, gauss, unif, jitter, shift = (t.to(x.device).float().contiguous() for t in draws)
        if mask:
            mk, mt = (t.float().contiguous() for t in mask_draws)
        ptr = lambda t: t.data_ptr() if t is not None else None   # noqa: E731
        lib = _lib.load()
        f32 = ctypes.c_float
        with torch.cuda.device(x.device):
            _lib.check(lib.mb_augment2d(
                xc.data_ptr(), cin, B, F, J, self.num_Kframes, int(noise), int(mask), ptr(sel), ptr(gauss), ptr(unif),
                ptr(jitter), ptr(shift), mean.data_ptr(), std.data_ptr(), weight.data_ptr(), f32(ur), f32(self.noise_std),
                f32(float(d["a"])), f32(float(d["b"])), f32(float(d["m"])), f32(float(d["s"])), ptr(mk), ptr(mt),
                f32(float(self.mask_ratio)), f32(float(self.mask_T_ratio)), out.data_ptr(),
                torch.cuda.current_stream(x.device).cuda_stream), "mb_augment2d")
        return out

    def _noise_draws(self, B, F, J):
        """the reference's draws, in its order (augmentation.py:43-47 then :24): all on the CPU generator"""
        K = self.num_Kframes
        sel = torch.rand((B, K, J, 1))
        gauss = torch.randn(B, K, J, 2)
        unif = torch.rand((B, K, J, 2))
        jitter = torch.randn(F, J, 2)
        shift = torch.randn(B, F, J)
        return sel, gauss, unif, jitter, shift

    @staticmethod
    def _mask_draws(x):
        N, T, J, _ = x.shape                                           # augmentation.py:71-72: drawn on x's device
        return (torch.rand(N, T, J, 1, dtype=x.dtype, device=x.device), torch.rand(1, T, 1, 1, dtype=x.dtype, device=x.device))

    def add_noise(self, motion_2d):
        B, F, J, _ = motion_2d.shape
        return self._launch(motion_2d, True, False, draws=self._noise_draws(B, F, J))

    def add_mask(self, x):
        return self._launch(x, False, True, mask_draws=self._mask_draws(x))

    def augment2D(self, motion_2d, mask=False, noise=False):
        if not (mask or noise):
            return motion_2d
        B, F, J, _ = motion_2d.shape
        draws = self._noise_draws(B, F, J) if noise else None
        mask_draws = self._mask_draws(motion_2d) if mask else None
        return self._launch(motion_2d, noise, mask, draws=draws, mask_draws=mask_draws)
