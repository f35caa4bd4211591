"""AdamW for the drop-in encoder (train.py:289 `optim.AdamW(...)`, train.py:206 `.step()`).

`AdamW` IS a `torch.optim.AdamW` -- same constructor arguments, `param_groups` (so the scripts' lr decay,
train.py:360-363, and per-group learning rates, train_action.py:143-147, keep working), same `state_dict()` layout
(`step` / `exp_avg` / `exp_avg_sq` per parameter, so the 'optimizer' entry of the reference's checkpoints,
train.py:46-54, loads and saves unchanged) -- but `step()` updates the encoder's parameter tensors through
`mb_adamw_step`: grouped launches over the handle's parameter table (6 kernels for 260 tensors) and an invalidation of
the packed tensor-core operands, which the next forward rebuilds in 3 grouped launches (was: 81).  Parameters that do
not belong to the encoder (task heads) are stepped by the inherited torch implementation in the same call.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class AdamW(torch.optim.AdamW):
    def __init__(self, encoder, params=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **kw):
        """encoder: the `motionbert_b200.DSTformer` whose tensors are stepped natively (unwrap DataParallel / ActionNet
        yourself: `model.module.backbone`).  params: iterable of tensors or param-group dicts exactly as for
        torch.optim.AdamW (default: the encoder's trainable parameters)."""
        if params is None:
            params = [p for p in encoder.parameters() if p.requires_grad]
        kw.pop("fused", None)
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)
        self._encoder = encoder

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        enc = self._encoder
        ordered = enc._ordered_params()
        slot = {id(p): i for i, p in enumerate(ordered) if p is not None}
        lib = enc._lib_loader()
        stashed = []
        for group in self.param_groups:
            if group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("motionbert_b200.optim.AdamW: amsgrad / maximize are not supported")
            native = [p for p in group["params"] if id(p) in slot and p.grad is not None]
            if not native:
                continue
            dev = native[0].device
            for p in native:
                if p.grad.is_sparse or p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                    raise NotImplementedError("native AdamW needs dense fp32 contiguous parameters on one device")
                st = self.state[p]
                if len(st) == 0:                                   # same lazy state as torch.optim.AdamW
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = {int(self.state[p]["step"]) for p in native}
            if len(steps) != 1:
                raise NotImplementedError("native AdamW: parameters of one group must share their step count")
            t = steps.pop() + 1
            n = len(ordered)
            vp = ctypes.c_void_p
            pp, gp, mp, sp = (vp * n)(), (vp * n)(), (vp * n)(), (vp * n)()
            active = bytearray(n)
            keep = []
            for p in native:
                i = slot[id(p)]
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(g)
                st = self.state[p]
                pp[i], gp[i], mp[i], sp[i] = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                active[i] = 1
            hstate = enc._state_for(dev, enc.train_math_mode)
            b1, b2 = group["betas"]
            f32 = ctypes.c_float
            with torch.cuda.device(dev):
                _lib.check(lib.mb_adamw_step(hstate.handle, pp, gp, mp, sp, bytes(active), t, f32(float(group["lr"])), f32(b1),
                                             f32(b2), f32(group["eps"]), f32(group["weight_decay"]),
                                             torch.cuda.current_stream(dev).cuda_stream), "mb_adamw_step")
            for p in native:
                self.state[p]["step"] += 1
                stashed.append((p, p.grad))
                p.grad = None                                       # so that the inherited step below skips it
        enc.invalidate_packed()                                     # the kernels wrote the weights behind autograd's back
        try:
            if any(p.grad is not None for g in self.param_groups for p in g["params"]):
                super().step()                                      # tensors outside the encoder (task heads)
        finally:
            for p, g in stashed:
                p.grad = g
        return loss
