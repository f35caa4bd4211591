"""Drop-in replacement for `lib.model.DSTformer.DSTformer` (reference: lib/model/DSTformer.py:269-361).

Same constructor signature, same parameter tree (260 tensors, identical names / shapes / init RNG
consumption), same `forward(x, return_rep=False)` / `get_representation(x)` / `get_classifier()` /
`reset_classifier()` surface -- so `lib/utils/learning.py::load_backbone`, `ActionNet`, `MeshRegressor`,
`nn.DataParallel`, `load_state_dict(strict=True)` and `partial_train_layers` keep working unchanged --
but the forward is ONE call into the sm_100a CUDA library through the C ABI (`mb_forward`; under autograd
`mb_forward_train`, with `loss.backward()` running `mb_backward`).

The sub-modules below (`nn.Linear`, `nn.LayerNorm`) are parameter containers only; they are never
called.  There is no CPU and no PyTorch-op fallback for the forward: a CPU tensor or a missing
library raises.
"""
from __future__ import annotations

import ctypes
import os
import math
from collections import OrderedDict

import threading

import torch
import torch.nn as nn

from . import _lib
from ._autograd import DSTformerFunction


class _Mlp(nn.Module):
    """Parameter container mirroring `MLP` (DSTformer.py:69-77): fc1, act, fc2, drop."""

    def __init__(self, dim, hidden, drop):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop = nn.Dropout(drop)


class _Attention(nn.Module):
    """Parameter container mirroring `Attention` (DSTformer.py:88-107): proj is built BEFORE qkv."""

    def __init__(self, dim, num_heads, qkv_bias, qk_scale, attn_drop, proj_drop, st_mode):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.mode = st_mode
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj_drop = nn.Dropout(proj_drop)


class _Block(nn.Module):
    """Parameter container mirroring `Block` (DSTformer.py:216-235)."""

    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path, norm_layer,
                 st_mode):
        super().__init__()
        self.st_mode = st_mode
        self.norm1_s = norm_layer(dim)
        self.norm1_t = norm_layer(dim)
        self.attn_s = _Attention(dim, num_heads, qkv_bias, qk_scale, attn_drop, drop, "spatial")
        self.attn_t = _Attention(dim, num_heads, qkv_bias, qk_scale, attn_drop, drop, "temporal")
        self.drop_path_rate = float(drop_path)
        self.drop_path = nn.Identity()      # DropPath itself is applied inside the kernels' residual epilogue
        self.norm2_s = norm_layer(dim)
        self.norm2_t = norm_layer(dim)
        hidden = int(dim * mlp_ratio)
        self.mlp_s = _Mlp(dim, hidden, drop)
        self.mlp_t = _Mlp(dim, hidden, drop)


class _DeviceState:
    """Per-device host state of one module: C handle, packed weights, workspaces."""

    def __init__(self):
        self.handle = None
        self.lib = None
        self.packed = None
        self.pack_key = None
        self.workspaces = {}
        self.pinned = set()          # workspace keys baked into a captured CUDA graph (make_graphed)
        self.zeros = {}
        self.side_stream = None
        self.graphs = {}             # (B, F, return_rep, kernel_flags) -> auto-captured inference graph (DSTformer._auto_graph)

    def __del__(self):
        try:
            if self.handle is not None:
                self.lib.mb_destroy(self.handle)
        except Exception:
            pass


class DSTformer(nn.Module):
    def __init__(self, dim_in=3, dim_out=3, dim_feat=256, dim_rep=512,
                 depth=5, num_heads=8, mlp_ratio=4,
                 num_joints=17, maxlen=243,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, att_fuse=True):
        super().__init__()
        if not dim_rep or dim_out <= 0:
            raise NotImplementedError("motionbert_b200.DSTformer needs dim_rep > 0 and dim_out > 0 "
                                      "(the only configuration lib/utils/learning.py:83-85 builds)")
        if not qkv_bias:
            raise NotImplementedError("qkv_bias=False is not supported (the factory always uses the default True)")
        self.dim_in = dim_in
        self.dim_out = dim_out
        self.dim_feat = dim_feat
        self.dim_rep = dim_rep
        self.depth = depth
        self.num_heads = num_heads
        self.num_joints = num_joints
        self.maxlen = maxlen
        self.hidden = int(dim_feat * mlp_ratio)
        self.qk_scale = qk_scale
        self.drop_rate = float(drop_rate)
        self.attn_drop_rate = float(attn_drop_rate)
        # ---- construction order below follows DSTformer.py:276-311 so that the RNG stream (and therefore
        # ---- seed-for-seed initial weights) is identical to the reference
        self.joints_embed = nn.Linear(dim_in, dim_feat)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth)]       # :279
        self.blocks_st = nn.ModuleList([
            _Block(dim_feat, num_heads, mlp_ratio, qkv_bias, qk_scale, drop_rate, attn_drop_rate, dpr[i], norm_layer,
                   "stage_st") for i in range(depth)])
        self.blocks_ts = nn.ModuleList([
            _Block(dim_feat, num_heads, mlp_ratio, qkv_bias, qk_scale, drop_rate, attn_drop_rate, dpr[i], norm_layer,
                   "stage_ts") for i in range(depth)])
        self.norm = norm_layer(dim_feat)
        self.pre_logits = nn.Sequential(OrderedDict([("fc", nn.Linear(dim_feat, dim_rep)), ("act", nn.Tanh())]))
        self.head = nn.Linear(dim_rep, dim_out)
        self.temp_embed = nn.Parameter(torch.zeros(1, maxlen, 1, dim_feat))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_joints, dim_feat))
        nn.init.trunc_normal_(self.temp_embed, std=.02)                          # :303
        nn.init.trunc_normal_(self.pos_embed, std=.02)                           # :304
        self.apply(self._init_weights)                                           # :305
        self.att_fuse = att_fuse
        if self.att_fuse:                                                        # :306-311
            self.ts_attn = nn.ModuleList([nn.Linear(dim_feat * 2, 2) for _ in range(depth)])
            for i in range(depth):
                self.ts_attn[i].weight.data.fill_(0)
                self.ts_attn[i].bias.data.fill_(0.5)
        if not isinstance(self.norm, nn.LayerNorm):
            raise NotImplementedError("norm_layer must build nn.LayerNorm (learning.py:84 passes "
                                      "partial(nn.LayerNorm, eps=1e-6))")
        self.eps = float(self.norm.eps)
        # arithmetic of the GEMM-shaped work (include/motionbert_b200.h MbMath): inference runs F16C (fp16 pass + one
        # e5m2 compensation pass, fp32 parity at 2 pass-equivalents); a forward that needs gradients runs BF16X3 (its
        # backward kernels consume bf16 operands).  set_math_mode() overrides both.
        self.math_mode = _lib.MB_MATH_F16C
        self.train_math_mode = _lib.MB_MATH_BF16X3
        self._kernel_flags = 0
        # forwards of at most this many tokens (B*F*J) are CUDA-graphed automatically in inference (0 / MB_AUTO_GRAPH=0: never)
        self.auto_graph_max_tokens = 0 if os.environ.get("MB_AUTO_GRAPH", "1") == "0" else 16384
        self._lib_loader = _lib.load               # tests may switch to the test twin (use_test_library)
        # shared (by reference) between nn.DataParallel replicas: keyed by device index
        self._dev_state = {}
        self._grad_sync = None                     # enable_gradient_allreduce(): {'group', 'world'}

    # ------------------------------------------------------------------ reference API surface
    def _init_weights(self, m):                                                  # DSTformer.py:313-320
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_classifier(self):                                                    # :322-323
        return self.head

    def reset_classifier(self, dim_out, global_pool=''):                         # :325-327 (quirk kept: dim_feat)
        self.dim_out = dim_out
        self.head = nn.Linear(self.dim_feat, dim_out) if dim_out > 0 else nn.Identity()
        self._dev_state.clear()

    def get_representation(self, x):                                             # :360-361
        return self.forward(x, return_rep=True)

    def get_representation_pooled(self, x):
        """mean over the frames of `get_representation(x)`: (B, F, J, 3) -> (B, J, dim_rep), what the action heads consume
        (lib/model/model_action.py:20-21).  Without autograd the tail GEMM's epilogue accumulates the mean and the
        (B, F, J, dim_rep) representation is never written (`mb_forward_pooled`); under autograd it is the plain mean."""
        if torch.is_grad_enabled() and (x.requires_grad or any(
                p is not None and p.requires_grad for p in self._ordered_params())):
            return self.get_representation(x).mean(dim=1)
        x = self._check_input(x)
        return self._launch(x, False, False, self._drop_path_scale(x.shape[0], x.shape[1], x.device), pooled=True)

    def set_math_mode(self, mode: str):
        """'f16c' (default: fp32 parity, fp16 pass + e5m2 compensation pass; gradient forwards use bf16x3),
        'bf16x3' (fp32 parity, 3 bf16 passes, inference and training) or 'bf16' (1 pass, inference and training)."""
        inf, trn = {"f16c": (_lib.MB_MATH_F16C, _lib.MB_MATH_BF16X3),
                    "bf16x3": (_lib.MB_MATH_BF16X3, _lib.MB_MATH_BF16X3),
                    "bf16": (_lib.MB_MATH_BF16, _lib.MB_MATH_BF16)}[mode]
        self.math_mode, self.train_math_mode = inf, trn
        self._dev_state.clear()
        return self

    def use_test_library(self):
        """TESTS ONLY: route this module through libmotionbert_b200_test.so (same source + the CUDA-core / first-generation
        reference kernels behind the MB_FLAG_REF_* / MB_FLAG_GEMM_1CTA test flags)."""
        self._lib_loader = _lib.load_test
        self._dev_state.clear()
        return self

    def invalidate_packed(self):
        """Force a weight re-pack at the next call.  Needed only after writes that bypass the version counter of the
        parameters (`p.data.copy_()`, EMA through `.data`): the packed-weight cache is keyed by (data_ptr, _version)."""
        for st in self._dev_state.values():
            st.pack_key = None
        return self

    # ------------------------------------------------------------------ host plumbing
    def _ordered_params(self):
        """The 260 tensors in the order mb_param_info() reports (== state_dict order of the reference)."""
        ps = [self.temp_embed, self.pos_embed, self.joints_embed.weight, self.joints_embed.bias]
        for blocks in (self.blocks_st, self.blocks_ts):
            for b in blocks:
                ps += [b.norm1_s.weight, b.norm1_s.bias, b.norm1_t.weight, b.norm1_t.bias]
                for a in (b.attn_s, b.attn_t):
                    ps += [a.proj.weight, a.proj.bias, a.qkv.weight, a.qkv.bias]
                ps += [b.norm2_s.weight, b.norm2_s.bias, b.norm2_t.weight, b.norm2_t.bias]
                for m in (b.mlp_s, b.mlp_t):
                    ps += [m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias]
        ps += [self.norm.weight, self.norm.bias, self.pre_logits.fc.weight, self.pre_logits.fc.bias,
               self.head.weight if isinstance(self.head, nn.Linear) else None,
               self.head.bias if isinstance(self.head, nn.Linear) else None]
        for i in range(self.depth):
            if self.att_fuse:
                ps += [self.ts_attn[i].weight, self.ts_attn[i].bias]
            else:
                ps += [None, None]       # zero logits -> alpha = (0.5, 0.5) == (x_st + x_ts) * 0.5 (:351)
        return ps

    def _state_for(self, device: torch.device, math: int = None) -> _DeviceState:
        """Host state (C handle, packed weights, workspaces) of one (device, math mode)."""
        if math is None:
            math = self.math_mode
        key = (device.index if device.index is not None else torch.cuda.current_device(), math)
        st = self._dev_state.get(key)
        if st is None:
            st = _DeviceState()
            lib = self._lib_loader()
            if not isinstance(self.head, nn.Linear) or self.head.in_features != self.dim_rep:
                raise NotImplementedError("head must be Linear(dim_rep, dim_out) for the fused tail")
            desc = _lib.MbDesc(self.dim_in, self.dim_out, self.dim_feat, self.dim_rep, self.depth, self.num_heads,
                               self.hidden, self.num_joints, self.maxlen, self.eps,
                               float(self.qk_scale) if self.qk_scale else 0.0, math)
            h = ctypes.c_void_p()
            _lib.check(lib.mb_create(ctypes.byref(desc), ctypes.byref(h)), "mb_create", lib)
            st.handle = h
            st.lib = lib
            n = _lib.check(lib.mb_param_count(h), "mb_param_count", lib)
            st.numels = []
            for i in range(n):
                ne = ctypes.c_int64()
                _lib.check(lib.mb_param_info(h, i, None, 0, ctypes.byref(ne)), "mb_param_info", lib)
                st.numels.append(ne.value)
            nb = ctypes.c_size_t()
            _lib.check(lib.mb_packed_bytes(h, ctypes.byref(nb)), "mb_packed_bytes", lib)
            st.packed = torch.empty(nb.value + 1024, dtype=torch.uint8, device=device)
            self._dev_state[key] = st
        return st

    @staticmethod
    def _evict_workspaces(st: _DeviceState):
        """Keep at most 4 inference workspaces per state; workspaces pinned by a captured CUDA graph are never dropped."""
        if len(st.workspaces) >= 4:
            for k in [k for k in st.workspaces if k not in st.pinned]:
                del st.workspaces[k]

    @staticmethod
    def _aligned_ptr(t: torch.Tensor) -> int:
        return (t.data_ptr() + 1023) // 1024 * 1024

    def _ensure_packed(self, st: _DeviceState, device: torch.device, stream_ptr: int):
        params = self._ordered_params()
        key = tuple((p.data_ptr(), p._version) if p is not None else (0, 0) for p in params)
        if key == st.pack_key:
            return
        lib = self._lib_loader()
        ptrs = (ctypes.c_void_p * len(params))()
        keep = []
        for i, p in enumerate(params):
            if p is None:
                z = st.zeros.get(st.numels[i])
                if z is None:
                    z = torch.zeros(st.numels[i], dtype=torch.float32, device=device)
                    st.zeros[st.numels[i]] = z
                t = z
            else:
                if p.device != device:
                    raise RuntimeError(f"parameter on {p.device} but input on {device}")
                t = p.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                    keep.append(t)
                if t.data_ptr() & 15:
                    # nn.DataParallel replicas hold their parameters as slices of ONE coalesced broadcast buffer
                    # (comm.broadcast_coalesced): everything after the 3-element head.bias sits at 4-byte alignment.
                    # The kernels read parameters with 16-byte accesses: hand them an aligned copy.
                    t = t.clone()
                    keep.append(t)
                if t.numel() != st.numels[i]:
                    raise RuntimeError(f"parameter {i} has {t.numel()} elements, library expects {st.numels[i]}")
            ptrs[i] = t.data_ptr()
        _lib.check(lib.mb_pack_weights(st.handle, ptrs, self._aligned_ptr(st.packed), stream_ptr), "mb_pack_weights", lib)
        st.pack_key = key
        st.keep = keep

    def _drop_path_scale(self, B, F, device):
        """Per-frame DropPath factors (lib/model/drop.py:17-32): floor(keep + U[0,1)) / keep, one independent
        draw per residual sublayer (4 per Block), only in training mode with rate > 0."""
        rates = [b.drop_path_rate for b in self.blocks_st]
        if not self.training or max(rates) <= 0.0:
            return None
        rows = []
        for i in range(self.depth):
            keep = 1.0 - rates[i]
            for _ in range(8):      # blocks_st[i] x4 then blocks_ts[i] x4 (mb_forward's sublayer order)
                if rates[i] <= 0.0:
                    rows.append(torch.ones(B * F, dtype=torch.float32, device=device))
                else:
                    r = keep + torch.rand(B * F, dtype=torch.float32, device=device)
                    rows.append(r.floor_() / keep)
        return torch.stack(rows).contiguous()

    def _launch(self, x: torch.Tensor, want_out: bool, want_rep: bool, dp_scale, pooled: bool = False):
        """One mb_forward (or mb_forward_pooled) call on the current stream of x.device.  Returns (out, rep), or the
        pooled representation."""
        device = x.device
        B, F, J, _ = x.shape
        lib = self._lib_loader()
        with torch.cuda.device(device):
            st = self._state_for(device)
            stream_ptr = torch.cuda.current_stream(device).cuda_stream
            self._ensure_packed(st, device, stream_ptr)
            ws = st.workspaces.get((B, F))
            if ws is None:
                nb = ctypes.c_size_t()
                _lib.check(lib.mb_workspace_bytes(st.handle, B, F, ctypes.byref(nb)), "mb_workspace_bytes", lib)
                self._evict_workspaces(st)
                ws = torch.empty(nb.value + 1024, dtype=torch.uint8, device=device)
                st.workspaces[(B, F)] = ws
            if pooled:
                if dp_scale is not None:
                    raise NotImplementedError("get_representation_pooled with DropPath active (training mode)")
                pool = torch.empty(B, J, self.dim_rep, dtype=torch.float32, device=device)
                _lib.check(lib.mb_forward_pooled(st.handle, self._aligned_ptr(st.packed), x.data_ptr(), pool.data_ptr(),
                                                 self._aligned_ptr(ws), ws.numel() - 1024, B, F, self._kernel_flags,
                                                 stream_ptr), "mb_forward_pooled", lib)
                return pool
            out = torch.empty(B, F, J, self.dim_out, dtype=torch.float32, device=device) if want_out else None
            rep = torch.empty(B, F, J, self.dim_rep, dtype=torch.float32, device=device) if want_rep else None
            _lib.check(lib.mb_forward(
                st.handle, self._aligned_ptr(st.packed), x.data_ptr(),
                out.data_ptr() if out is not None else None, rep.data_ptr() if rep is not None else None,
                dp_scale.data_ptr() if dp_scale is not None else None,
                self._aligned_ptr(ws), ws.numel() - 1024, B, F, self._kernel_flags, stream_ptr), "mb_forward", lib)
        return out, rep

    # ------------------------------------------------------------------ training (mb_forward_train / mb_backward)
    def _check_native_backward(self, x: torch.Tensor):
        """The hand-written backward needs fp32 contiguous parameters on x's device and dim_out <= 8 (every
        configuration the reference's training scripts build).  Anything else raises: there is no PyTorch-op fallback."""
        if self.dim_out > 8:
            raise NotImplementedError(f"training with dim_out={self.dim_out} > 8 is not supported by the native backward "
                                      "(head_bwd_kernel keeps the head gradient in registers)")
        for p in self._ordered_params():
            if p is not None and (p.dtype != torch.float32 or not p.is_contiguous() or p.device != x.device):
                raise NotImplementedError("the native backward needs fp32, contiguous parameters on the input's device "
                                          f"(got {p.dtype}, contiguous={p.is_contiguous()}, {p.device} vs {x.device})")

    def _head_param_slots(self):
        """Positions of head.weight / head.bias among the non-None tensors of `_ordered_params()`."""
        live = [p for p in self._ordered_params() if p is not None]
        ids = {id(self.head.weight), id(self.head.bias)} if isinstance(self.head, nn.Linear) else set()
        return tuple(i for i, p in enumerate(live) if id(p) in ids)

    def _launch_train(self, x: torch.Tensor, want_out: bool, dp_scale=None):
        """mb_forward_train on the current stream: returns (out, rep, saved) with `saved` the activation region."""
        device = x.device
        B, F, J, _ = x.shape
        lib = self._lib_loader()
        with torch.cuda.device(device):
            st = self._state_for(device, self.train_math_mode)
            stream_ptr = torch.cuda.current_stream(device).cuda_stream
            self._ensure_packed(st, device, stream_ptr)
            ws = st.workspaces.get((B, F))
            if ws is None:
                nb = ctypes.c_size_t()
                _lib.check(lib.mb_workspace_bytes(st.handle, B, F, ctypes.byref(nb)), "mb_workspace_bytes", lib)
                self._evict_workspaces(st)
                ws = torch.empty(nb.value + 1024, dtype=torch.uint8, device=device)
                st.workspaces[(B, F)] = ws
            nb = ctypes.c_size_t()
            _lib.check(lib.mb_saved_bytes(st.handle, B, F, ctypes.byref(nb)), "mb_saved_bytes", lib)
            saved = torch.empty(nb.value + 1024, dtype=torch.uint8, device=device)   # one per forward call
            out = torch.empty(B, F, J, self.dim_out, dtype=torch.float32, device=device) if want_out else None
            rep = torch.empty(B, F, J, self.dim_rep, dtype=torch.float32, device=device)
            _lib.check(lib.mb_forward_train(
                st.handle, self._aligned_ptr(st.packed), x.data_ptr(), out.data_ptr() if out is not None else None,
                rep.data_ptr(), dp_scale.data_ptr() if dp_scale is not None else None, self._aligned_ptr(saved),
                saved.numel() - 1024, self._aligned_ptr(ws), ws.numel() - 1024, B, F, self._kernel_flags, stream_ptr),
                "mb_forward_train", lib)
        return out, rep, saved

    def _param_phases(self):
        """Backward phase of every tensor of `_ordered_params()`: 0 tail, 1 + (depth-1-i) depth i, depth + 1 embed
        (the order mb_backward finishes their gradients in; include/motionbert_b200.h `phase_events`)."""
        d = self.depth
        ph = [d + 1] * 4
        for _stream in range(2):
            for i in range(d):
                ph += [1 + (d - 1 - i)] * 24
        ph += [0] * 6
        for i in range(d):
            ph += [1 + (d - 1 - i)] * 2
        return ph

    def enable_gradient_allreduce(self, group=None, enabled: bool = True):
        """Data-parallel training without a wrapper: every native backward averages
        the parameter gradients over the ranks of `group`, phase by phase (tail, depth d-1 ... 0, embed) on a side
        stream, overlapped with the rest of the backward.  Needs an initialised torch.distributed process group (NCCL).
        Replaces `nn.DataParallel`'s reduce (train.py:258) / an external DistributedDataParallel wrapper."""
        if not enabled:
            self._grad_sync = None
            return self
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("enable_gradient_allreduce needs an initialised torch.distributed process group")
        self._grad_sync = {"group": group, "world": dist.get_world_size(group)}
        return self

    def _launch_backward(self, x, rep, saved, d_out, d_rep, dp_scale=None, want_dx=False):
        """mb_backward on the current stream: returns (gradients of the non-None tensors of `_ordered_params()`, in that
        order; d_x or None).  `att_fuse=False`: the missing fusion-head slots are fed zero weights (alpha = 0.5 / 0.5, which
        IS DSTformer.py:351) and their gradients land in a scratch tensor that is dropped."""
        device = x.device
        B, F, J, _ = x.shape
        lib = self._lib_loader()
        real = self._ordered_params()
        st0 = self._state_for(device, self.train_math_mode)
        params = []
        for i, p in enumerate(real):
            if p is None:
                z = st0.zeros.get(st0.numels[i])
                if z is None:
                    z = st0.zeros[st0.numels[i]] = torch.zeros(st0.numels[i], dtype=torch.float32, device=device)
                p = z
            params.append(p)
        with torch.cuda.device(device):
            st = self._state_for(device, self.train_math_mode)
            stream_ptr = torch.cuda.current_stream(device).cuda_stream
            self._ensure_packed(st, device, stream_ptr)      # no-op unless the weights changed since the forward
            bws = st.workspaces.get(("bwd", B, F))
            if bws is None:
                nb = ctypes.c_size_t()
                _lib.check(lib.mb_backward_workspace_bytes(st.handle, B, F, ctypes.byref(nb)), "mb_backward_workspace_bytes", lib)
                bws = torch.empty(nb.value + 1024, dtype=torch.uint8, device=device)
                st.workspaces[("bwd", B, F)] = bws
            # ONE zero-filled flat bucket, laid out phase by phase (tail, depth d-1 ... 0, embed) in the order the backward
            # finishes them, every tensor 256-byte aligned: the DDP bucket of SURVEY.md section 8e
            phases = self._param_phases()
            nph = self.depth + 2
            sizes = [(p.numel() + 63) // 64 * 64 for p in params]
            order = sorted(range(len(params)), key=lambda i: phases[i])
            offs, bounds, off = [0] * len(params), [0] * (nph + 1), 0
            for i in order:
                offs[i] = off
                off += sizes[i]
                bounds[phases[i] + 1] = off
            flat = torch.zeros(off, dtype=torch.float32, device=device)
            grads = [flat[offs[i]:offs[i] + p.numel()].view(p.shape) for i, p in enumerate(params)]
            # (16-byte aligned parameter storage: see _ensure_packed -- replicas of nn.DataParallel are not)
            params = [p if not (p.data_ptr() & 15) else p.detach().clone() for p in params]
            pp = (ctypes.c_void_p * len(params))(*[p.data_ptr() for p in params])
            gp = (ctypes.c_void_p * len(params))(*[g.data_ptr() for g in grads])
            d_x = torch.empty_like(x) if want_dx else None
            sync = self._grad_sync if self._grad_sync is not None and self._grad_sync["world"] > 1 else None
            events, ev_ptrs = None, None
            if sync is not None:
                cur = torch.cuda.current_stream(device)
                events = [torch.cuda.Event() for _ in range(nph)]
                for e in events:
                    e.record(cur)                       # materialise the cudaEvent_t; the library re-records it
                ev_ptrs = (ctypes.c_void_p * nph)(*[e.cuda_event for e in events])
            _lib.check(lib.mb_backward(
                st.handle, self._aligned_ptr(st.packed), pp, x.data_ptr(), rep.data_ptr(), self._aligned_ptr(saved),
                saved.numel() - 1024, dp_scale.data_ptr() if dp_scale is not None else None,
                d_out.data_ptr() if d_out is not None else None, d_rep.data_ptr() if d_rep is not None else None, gp,
                d_x.data_ptr() if d_x is not None else None, self._aligned_ptr(bws), bws.numel() - 1024,
                B, F, ev_ptrs, stream_ptr), "mb_backward", lib)
            if sync is not None:
                # data-parallel exchange overlapped with the backward: phase k is summed over the ranks on a side stream
                # as soon as its event fires, while the kernels of phase k+1 keep the SMs busy on the main stream
                import torch.distributed as dist
                side = st.side_stream
                if side is None:
                    side = st.side_stream = torch.cuda.Stream(device=device)
                flat.record_stream(side)
                with torch.cuda.stream(side):
                    for k in range(nph):
                        if bounds[k + 1] > bounds[k]:
                            side.wait_event(events[k])
                            seg = flat[bounds[k]:bounds[k + 1]]
                            dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=sync["group"])
                            seg.div_(sync["world"])
                torch.cuda.current_stream(device).wait_stream(side)
        return [g for g, p in zip(grads, real) if p is not None], d_x

    def make_graphed(self, B: int, F: int, return_rep: bool = False):
        """CUDA-graph the inference forward for a fixed (B, F): returns `run(x) -> out` that copies x into a static
        device buffer, replays the captured forward and returns the static output tensor (valid until the next call).
        For latency-bound shapes (infer_wild: B=1 clips).  The graph bakes in device pointers of the workspace and of the
        packed weights: both are pinned for the lifetime of `run` (held by the closure, exempt from workspace eviction),
        and `run` raises if the parameters changed or the module state was reset since the capture (re-capture then)."""
        dev = next(self.parameters()).device
        static_x = torch.zeros(B, F, self.num_joints, self.dim_in, dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(2):                      # warm-up: handle, packed weights, workspace, tensor maps
                self.forward(static_x, return_rep)
        torch.cuda.current_stream(dev).wait_stream(side)
        st = self._state_for(dev)
        st.pinned.add((B, F))
        keep = (st, st.workspaces[(B, F)], st.packed)           # the buffers the captured launches point into
        pack_key = st.pack_key
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            static_out = self.forward(static_x, return_rep)

        def run(x):
            cur = self._dev_state.get((dev.index if dev.index is not None else torch.cuda.current_device(), self.math_mode))
            params = self._ordered_params()
            key = tuple((p.data_ptr(), p._version) if p is not None else (0, 0) for p in params)
            if cur is not keep[0] or key != pack_key:
                raise RuntimeError("make_graphed: parameters or math mode changed since the capture; call make_graphed again")
            static_x.copy_(x, non_blocking=True)
            graph.replay()
            return static_out
        run.graph = graph
        run._keep = keep
        return run

    # ------------------------------------------------------------------ forward (DSTformer.py:329-358)
    def _check_input(self, x):
        if x.dim() != 4:
            raise ValueError(f"expected (B, F, J, C) input, got {tuple(x.shape)}")
        B, F, J, Cin = x.shape
        if J != self.num_joints or Cin != self.dim_in:
            raise RuntimeError(f"input (…, {J}, {Cin}) does not match num_joints={self.num_joints}, dim_in={self.dim_in}")
        if F > self.maxlen:
            raise RuntimeError(f"sequence length {F} exceeds maxlen {self.maxlen} (temp_embed, DSTformer.py:336)")
        if not x.is_cuda:
            raise RuntimeError("motionbert_b200.DSTformer runs on sm_100a CUDA devices only: there is no CPU fallback "
                               "(the reference's CPU path is timed separately as a baseline)")
        if self.training and (self.drop_rate > 0 or self.attn_drop_rate > 0):
            raise NotImplementedError("dropout / attention dropout > 0 in training mode is not implemented "
                                      "(all shipped configs use 0; DropPath is supported)")
        return x.detach().float().contiguous() if not x.requires_grad else x.float().contiguous()

    def forward(self, x, return_rep=False):
        x = self._check_input(x)
        B, F, J, Cin = x.shape
        dp_scale = self._drop_path_scale(B, F, x.device)
        # (not self.parameters(): nn.DataParallel replicas carry their parameters as plain attributes)
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(
            p is not None and p.requires_grad for p in self._ordered_params()))
        if needs_grad:
            params = [p for p in self._ordered_params() if p is not None]
            return DSTformerFunction.apply(self, x, bool(return_rep), dp_scale, *params)
        if dp_scale is None and self.auto_graph_max_tokens > 0 and B * F * J <= self.auto_graph_max_tokens:
            res = self._auto_graph(x, bool(return_rep))
            if res is not None:
                return res
        out, rep = self._launch(x, not return_rep, bool(return_rep), dp_scale)
        return rep if return_rep else out

    # ------------------------------------------------------------------ launch-bound shapes: automatic CUDA-graph replay
    AUTO_GRAPH_AFTER = 2         # eager calls of one (B, F, output kind) before its forward is captured
    AUTO_GRAPH_MAX = 8           # captured shapes per device

    def _auto_graph(self, x, return_rep):
        """Inference forwards of small clips (B*F*J <= auto_graph_max_tokens, e.g. infer_wild's B=1 windows): from the third
        call of a shape on, the forward is captured into a CUDA graph and replayed (one launch instead of ~90; measured
        ~10 % lower latency, profiles/r02i_latency_auto_graph.log -- these shapes are bound by the per-kernel pipeline
        fill/drain more than by the launches).  Returns None when this call has to run eagerly.  The graph bakes in the
        pointers of the workspace and of the packed weights: both are pinned by the cache entry, and an entry dies as soon
        as the parameters change (`pack_key`).  The result is a fresh tensor (a copy of the graph's static output).  A
        failed capture disables graphing for that shape (the call and all later ones run eagerly)."""
        if torch.cuda.is_current_stream_capturing() or threading.current_thread() is not threading.main_thread():
            return None
        device = x.device
        B, F = int(x.shape[0]), int(x.shape[1])
        with torch.cuda.device(device):
            st = self._state_for(device)
            self._ensure_packed(st, device, torch.cuda.current_stream(device).cuda_stream)
            key = (B, F, return_rep, self._kernel_flags)
            ent = st.graphs.get(key)
            if ent is not None and ent.get("graph") is not None and ent["pack_key"] != st.pack_key:
                st.pinned.discard((B, F))
                ent = None                                     # parameters changed: capture again after the warm-up calls
            if ent is None:
                if len(st.graphs) >= self.AUTO_GRAPH_MAX and key not in st.graphs:
                    return None
                st.graphs[key] = {"hits": 1, "graph": None}
                return None
            if ent.get("dead"):
                return None
            if ent["graph"] is None:
                ent["hits"] += 1
                if ent["hits"] <= self.AUTO_GRAPH_AFTER or (B, F) not in st.workspaces:
                    return None                                # (workspace evicted meanwhile: one more eager call)
                static_x = torch.empty_like(x)
                static_x.copy_(x)
                was_pinned = (B, F) in st.pinned
                st.pinned.add((B, F))
                graph = torch.cuda.CUDAGraph()
                try:
                    with torch.no_grad(), torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        out, rep = self._launch(static_x, not return_rep, return_rep, None)
                except Exception:
                    ent["dead"] = True                         # e.g. a capture-unsafe call elsewhere in the process
                    if not was_pinned:
                        st.pinned.discard((B, F))
                    return None
                ent.update(graph=graph, x=static_x, y=rep if return_rep else out, pack_key=st.pack_key,
                           keep=(st.workspaces[(B, F)], st.packed))
            ent["x"].copy_(x, non_blocking=True)
            ent["graph"].replay()
            return ent["y"].clone()
