#!/usr/bin/env python
"""bench.py -- sequences/sec of the DSTformer forward (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU path (oracle port), rank 0 only

A "step" is one forward of the hot path over one batch of synthetic 2D skeleton clips.  At N=1 the workload is
BASELINE config 2: DSTformer-base forward, B=256, T=243, 17 joints, fp32-parity arithmetic (BF16x3).  Multi-GPU:
the batch shards across ranks as independent sequences (weak scaling: B=256 per GPU), no data-path collective;
NCCL is only used for the barrier and the max-over-ranks time.

JSON keys follow the driver contract; `roofline` is the tcgen05 GEMM class (the dominant kernels), measured with
CUDA events around every launch in a separate profiled pass of the same steps.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}   # B200_PROFILING.md


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback (B200_PROFILING.md)"
    return d


MODELS = {"base": dict(dim_feat=512, mlp_ratio=2), "lite": dict(dim_feat=256, mlp_ratio=4)}


def flops_per_sequence(dim_feat, hidden, T, J=17, depth=5, dim_rep=512, dim_in=3, dim_out=3):
    """Algorithmic FLOPs of one forward (equals torch FlopCounterMode on the reference)."""
    C = dim_feat
    tok = depth * (4 * 8 * C * C + 4 * 4 * C * hidden + 2 * 4 * J * C + 2 * 4 * T * C + 8 * C) \
        + 2 * dim_in * C + 2 * C * dim_rep + 2 * dim_rep * dim_out
    return float(tok) * T * J


def gemm_flops_per_sequence(dim_feat, hidden, T, J=17, depth=5, dim_rep=512):
    """FLOPs of the work the tcgen05 GEMM kernel does (qkv/proj/fc1/fc2 of 20 sublayers + pre_logits)."""
    C = dim_feat
    tok = depth * (4 * 8 * C * C + 4 * 4 * C * hidden) + 2 * C * dim_rep
    return float(tok) * T * J


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                    str(self.index)], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for row in self.rows:
            try:
                sm.append(float(row[0])); mx.append(float(row[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), row[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(model, device, math):
    from functools import partial

    import torch.nn as nn

    from motionbert_b200 import DSTformer
    cfg = MODELS[model]
    torch.manual_seed(0)                      # reference init (random-init weights of that architecture)
    m = DSTformer(dim_in=3, dim_out=3, dim_feat=cfg["dim_feat"], dim_rep=512, depth=5, num_heads=8,
                  mlp_ratio=cfg["mlp_ratio"], norm_layer=partial(nn.LayerNorm, eps=1e-6), maxlen=243, num_joints=17)
    # move the S/T fusion and LayerNorm affine off their trivial init so no sub-path is constant
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for i in range(5):
            m.ts_attn[i].weight.normal_(0, 0.05, generator=g)
        for mod in m.modules():
            if isinstance(mod, nn.LayerNorm):
                mod.weight.add_(0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.add_(0.05 * torch.randn(mod.bias.shape, generator=g))
    m = m.to(device).eval()
    m.set_math_mode(math)
    return m


def synthetic_clips(B, T, seed):
    """x,y ~ U(-1,1), confidence ~ U(0,1), pinned host memory."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, T, 17, 3, generator=g)
    x[..., :2] = x[..., :2] * 2 - 1
    return x.pin_memory() if torch.cuda.is_available() else x


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_rate(model, T, budget_s=20.0, batch=2, max_iters=12):
    """The reference's own CPU forward, restated op-for-op with torch CPU ops (oracle/dstformer_torch_cpu.py; the
    reference itself is PyTorch-only Python and cannot travel to the GPU box).  Uses the host thread count that a
    short calibration finds fastest (torch's default = all cores is pathological on 100+-thread hosts)."""
    from oracle import dstformer_oracle as O
    from oracle import dstformer_torch_cpu as OT
    cfg = O.BASE if model == "base" else O.LITE
    cores = os.cpu_count() or 1
    P = {k: torch.from_numpy(v) for k, v in O.make_params(cfg, 0).items()}
    xs = torch.from_numpy(O.make_input(1, min(T, 81), cfg.num_joints, 2))
    best_t, best_n = None, cores
    for n in sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 16), min(cores, 8)}, reverse=True):
        torch.set_num_threads(n)
        OT.forward(P, xs, cfg.depth, cfg.num_heads, cfg.eps)
        t0 = time.perf_counter()
        OT.forward(P, xs, cfg.depth, cfg.num_heads, cfg.eps)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n
    torch.set_num_threads(best_n)
    x = torch.from_numpy(O.make_input(batch, T, cfg.num_joints, 1))
    OT.forward(P, x, cfg.depth, cfg.num_heads, cfg.eps)           # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < max_iters and (not times or (time.perf_counter() - t_start) < budget_s):
        t0 = time.perf_counter()
        OT.forward(P, x, cfg.depth, cfg.num_heads, cfg.eps)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"value": batch / med, "unit": "sequences/sec", "cores": best_n, "kind": "port",
            "sample": f"{len(times)} forwards of B={batch} x T={T} x 17 ({model}), median {med * 1e3:.0f} ms, "
                      f"torch {torch.__version__} CPU fp32, {best_n} of {cores} host threads (fastest in calibration)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    cb = cpu_reference_rate(args.model, args.frames, budget_s=min(120.0, 8.0 * steps), batch=2, max_iters=steps + args.warmup)
    line = {
        "impl": "reference", "metric": f"sequences/sec DSTformer-{args.model} fwd (Bx{args.frames}x17)",
        "value": cb["value"], "unit": "sequences/sec", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 2.0 / cb["value"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DSTformer-{args.model} forward, T={args.frames}, 17 joints, fp32; bounded CPU sample B=2 per step"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "sequences/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def _barrier(dist, device):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)


def train_record(args, device, world, rank, local_rank, dist, D, model_name="base", batch=128, frames=243, math="bf16",
                 steps=20, warmup=3, e2e_steps=0):
    """SURVEY.md 8d config 3 (N=1) / config 4 (N>1): one pretrain step = forward (saved residual stream) -> fused pretrain
    loss (mpjpe + 0.5 n_mpjpe + 20 velocity, train.py:178-191) -> native backward -> gradient all-reduce over the ranks
    (N>1, per depth, overlapped with the backward) -> fused AdamW -> weight re-pack at the next forward.
    Returns the record (identical on every rank: times are max over ranks)."""
    from motionbert_b200 import _lib
    from motionbert_b200.loss import pretrain_loss_3d
    model = build_model(model_name, device, math).train()
    cfg = MODELS[model_name]
    hidden = int(cfg["dim_feat"] * cfg["mlp_ratio"])
    B, T = batch, frames
    x_host = synthetic_clips(B, T, seed=1 + rank)
    gt_host = synthetic_clips(B, T, seed=1001 + rank)
    x_dev, gt_dev = x_host.to(device), gt_host.to(device)
    params = [p for p in model.parameters()]
    if world > 1:
        model.enable_gradient_allreduce()
    from motionbert_b200.optim import AdamW
    opt = AdamW(model, params, lr=1e-5, weight_decay=0.01)          # grouped native step + grouped re-pack (row f4)

    def step(x, gt):
        opt.zero_grad(set_to_none=True)
        pred = model(x)
        loss, _parts = pretrain_loss_3d(pred, gt, 0.5, 20.0)
        loss.backward()       # world > 1: gradients are averaged over the ranks inside the backward (phase by phase)
        opt.step()
        return loss

    for _ in range(warmup):
        step(x_dev, gt_dev)
    _barrier(dist, device)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(dist, device)
    ev0.record()
    for _ in range(steps):
        loss = step(x_dev, gt_dev)
    ev1.record()
    _barrier(dist, device)
    ms_per_step = D.max_over_ranks(ev0.elapsed_time(ev1), device) / steps
    sampler.stop_flag = True
    sampler.join(timeout=3)
    clocks = sampler.summary()
    last = float(loss.detach())
    e2e = None
    if e2e_steps > 0:
        # end to end: pinned host clip + target -> H2D, step, loss read back to the host every step
        _barrier(dist, device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            last = float(step(x_host.to(device, non_blocking=True), gt_host.to(device, non_blocking=True)).item())
        e1.record()
        _barrier(dist, device)
        e2e_ms = D.max_over_ranks(e0.elapsed_time(e1), device) / e2e_steps
        e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": "sequences/sec", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(2 * x_host.numel() * 4), "d2h_bytes_per_step": 4,
               "api": "DSTformer.forward -> loss.backward() -> optimizer.step() on pinned host clips"}
    peaks = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    step_flop = 3.0 * flops_per_sequence(cfg["dim_feat"], hidden, T) * B          # fwd + 2x bwd (recompute not counted)
    achieved = step_flop / (ms_per_step * 1e-3) / 1e12
    n_param = sum(p.numel() for p in params)
    lib = _lib.load()
    hnd = model._state_for(device, model.train_math_mode).handle
    launches_per_step = (_lib.check(lib.mb_forward_launch_count(hnd, 1, 0)) +
                         _lib.check(lib.mb_backward_launch_count(hnd, 0, 0)) + 2)
    rec = {
        "metric": f"sequences/sec DSTformer-{model_name} pretrain step fwd+bwd+AdamW (Bx{T}x17)",
        "config": f"SURVEY 8d config {'3' if world == 1 else '4'}: B={B} per GPU, T={T}, {math} forward, bf16 native backward, "
                  "fused pretrain loss, native grouped AdamW + grouped weight re-pack" + (", per-depth NCCL gradient all-reduce overlapped with the backward" if world > 1 else ""),
        "value": world * B / (ms_per_step * 1e-3), "unit": "sequences/sec", "n_gpus": world, "global_batch": world * B,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "last_loss": last, "clocks": clocks,
        "whole_step_tflops": achieved, "whole_step_frac": achieved / peak_tf,
        "frac_note": "3 x forward algorithmic FLOPs / step time / sustained bf16 peak (recompute, optimizer, all-reduce in the time only)",
        "grad_allreduce_bytes_per_step": n_param * 4 if world > 1 else 0,
        "gpu_launches_per_step": launches_per_step,
    }
    if e2e is not None:
        rec["e2e"] = e2e
    del opt, model, x_dev, gt_dev
    torch.cuda.empty_cache()
    return rec


def run_train(args, device, world, rank, local_rank, dist, D):
    """`--mode train`: the training step as the headline line (supplementary to the forward line)."""
    rec = train_record(args, device, world, rank, local_rank, dist, D, args.model, args.batch, args.frames, args.math,
                       steps=args.steps, warmup=args.warmup, e2e_steps=args.steps)
    peaks = load_peaks()
    line = {
        "metric": rec["metric"], "value": rec["value"], "unit": "sequences/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.math == "bf16" else "bf16x3 forward / bf16 backward", "data": "synthetic",
        "config": {"workload": rec["config"], "global_batch": world * args.batch, "seq_len": args.frames,
                   "parallelism": f"dp{world}", "l2": "activations >> 126 MB L2, no flush needed"},
        "clocks": rec["clocks"], "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches_per_step"] * args.steps,
        "roofline": {"bound": "tensor", "achieved": rec["whole_step_tflops"],
                     "peak": float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))), "unit": "TFLOP/s",
                     "frac": rec["whole_step_frac"], "traffic": None, "note": rec["frac_note"],
                     "peak_source": peaks["_source"] + " bf16_tflops_sustained"},
    }
    if rank == 0:
        print(json.dumps(line), flush=True)


def forward_rate(model, x_dev, steps, warmup, dist, device, D):
    """K forwards with the input resident in HBM, CUDA events on the launching stream, max over ranks -> ms per step."""
    with torch.no_grad():
        for _ in range(warmup):
            out = model(x_dev)
    _barrier(dist, device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(dist, device)
    with torch.no_grad():
        ev0.record()
        for _ in range(steps):
            out = model(x_dev)
        ev1.record()
    _barrier(dist, device)
    return D.max_over_ranks(ev0.elapsed_time(ev1), device) / steps, out


def parity_check(model, model_name, x_host, out_dev, device, n_check=3):
    """The output the bench timed, checked against the CPU oracle: sequences first / middle / last of the batch through
    `oracle/dstformer_torch_cpu.py` in float64 (op-for-op the reference's forward, pinned against the real reference by
    tests/golden); north-star bars: per-token relative error <= 1e-3, MPJPE <= 2e-4 units (0.1 mm at 500 mm / unit).
    Also re-runs the three sequences as their own batch for `rep` (the 512-d representation)."""
    from oracle import dstformer_oracle as O
    from oracle import dstformer_torch_cpu as OT
    cfg = O.BASE if model_name == "base" else O.LITE
    B = x_host.shape[0]
    idx = sorted({0, B // 2, B - 1})[:n_check]
    P64 = {k: v.detach().double().cpu() for k, v in model.state_dict().items()}
    xs = x_host[idx].double()
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    o_ref, r_ref = OT.forward(P64, xs, cfg.depth, cfg.num_heads, cfg.eps)
    got = out_dev[idx].double().cpu()
    with torch.no_grad():
        rep = model.get_representation(x_host[idx].to(device)).double().cpu()
    tok = (got - o_ref).norm(dim=-1) / o_ref.norm(dim=-1).clamp_min(1e-12)
    rtok = (rep - r_ref).norm(dim=-1) / r_ref.norm(dim=-1).clamp_min(1e-12)
    mp = float((got - o_ref).norm(dim=-1).mean())
    res = {"checked_sequences": idx, "of_batch": B, "out_tok_rel_mean": float(tok.mean()), "out_tok_rel_max": float(tok.max()),
           "rep_tok_rel_mean": float(rtok.mean()), "rep_tok_rel_max": float(rtok.max()), "mpjpe_units": mp,
           "mpjpe_mm_at_500mm_per_unit": mp * 500.0, "out_joint_norm_mean": float(o_ref.norm(dim=-1).mean()),
           "checker": "oracle/dstformer_torch_cpu.py float64 (CPU)", "bars": {"tok_rel": 1e-3, "mpjpe_units": 2e-4}}
    res["ok"] = bool(res["rep_tok_rel_max"] < 1e-3 and res["out_tok_rel_mean"] < 1e-3 and mp < 2e-4)
    return res


def gpu_eager_baseline(model, model_name, T, device, batch=32, iters=3):
    """The honest GPU comparator (SURVEY.md 8d / BASELINE.md 3): the reference's forward op for op -- the bit-exact torch
    port `oracle/dstformer_torch_cpu.py`, which is plain torch ops -- moved to the B200 in eager mode, fp32 and TF32-allowed.
    A reported baseline; nothing of it is on the product path."""
    from oracle import dstformer_oracle as O
    from oracle import dstformer_torch_cpu as OT
    cfg = O.BASE if model_name == "base" else O.LITE
    P = {k: v.detach().to(device) for k, v in model.state_dict().items()}
    x = synthetic_clips(batch, T, seed=5).to(device)
    res = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            for _ in range(2):
                OT.forward(P, x, cfg.depth, cfg.num_heads, cfg.eps)
            torch.cuda.synchronize(device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                OT.forward(P, x, cfg.depth, cfg.num_heads, cfg.eps)
            e1.record()
            torch.cuda.synchronize(device)
            ms = e0.elapsed_time(e1) / iters
            res[name] = {"value": batch / (ms * 1e-3), "unit": "sequences/sec", "ms_per_step": ms}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    res["what"] = (f"torch eager (cuBLAS / ATen kernels) port of lib/model/DSTformer.py forward on the same B200, B={batch}, T={T}, "
                   f"{iters} iterations after 2 warm-ups; torch {torch.__version__}")
    del P, x
    torch.cuda.empty_cache()
    return res


def gpu_eager_train_baseline(model, model_name, T, device, batch=16, iters=3):
    """SURVEY.md 8d, config 3's GPU comparator: the reference's op sequence (the torch port of lib/model/DSTformer.py) as a
    TRAINING step in torch eager on this device -- forward under torch.autocast(bf16) (and, second entry, plain fp32 with
    TF32 matmuls), an MPJPE-style loss, autograd backward, torch's fused AdamW.  A reported baseline at a batch that fits the
    eager activations (the reference materialises the (B, 8, 17, T, T) score tensors); nothing of it is on the product path."""
    from oracle import dstformer_oracle as O
    from oracle import dstformer_torch_cpu as OT
    cfg = O.BASE if model_name == "base" else O.LITE
    P = {k: v.detach().clone().to(device).requires_grad_(True) for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(list(P.values()), lr=1e-5, weight_decay=0.01, fused=(device.type == "cuda"))
    x = synthetic_clips(batch, T, seed=5).to(device)
    gt = synthetic_clips(batch, T, seed=6).to(device)
    res = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    fwd = OT.forward.__wrapped__          # the port itself (its public entry is wrapped in torch.no_grad for the CPU baseline)

    def step(autocast):
        opt.zero_grad(set_to_none=True)
        with torch.autocast(device.type, dtype=torch.bfloat16, enabled=autocast):
            out, _rep = fwd(P, x, cfg.depth, cfg.num_heads, cfg.eps)
        loss = (out.float() - gt).norm(dim=-1).mean()
        loss.backward()
        opt.step()
        return loss

    try:
        for name, autocast, tf32 in (("bf16_autocast", True, False), ("tf32", False, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            for _ in range(2):
                step(autocast)
            if device.type == "cuda":
                torch.cuda.synchronize(device)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    step(autocast)
                e1.record()
                torch.cuda.synchronize(device)
                ms = e0.elapsed_time(e1) / iters
            else:                                                   # (CPU: only used to exercise this function in the tests)
                t0 = time.perf_counter()
                for _ in range(iters):
                    step(autocast)
                ms = (time.perf_counter() - t0) * 1e3 / iters
            res[name] = {"value": batch / (ms * 1e-3), "unit": "sequences/sec", "ms_per_step": ms}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    res["what"] = (f"torch eager training step (port of lib/model/DSTformer.py forward, autograd backward, torch.optim.AdamW fused) "
                   f"on the same device, B={batch}, T={T}, {iters} iterations after 2 warm-ups; torch {torch.__version__}")
    del P, x, gt, opt
    if device.type == "cuda":
        torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="base", choices=["base", "lite"])
    ap.add_argument("--mode", default="forward", choices=["forward", "train"],
                    help="forward = BASELINE config 2 (the headline, with train / lite / eager sub-records); train = config 3/4 only")
    ap.add_argument("--batch", type=int, default=None, help="sequences per GPU per step (default 256 forward / 128 train)")
    ap.add_argument("--frames", type=int, default=243)
    ap.add_argument("--math", default=None, choices=["f16c", "bf16x3", "bf16"],
                    help="default f16c (fp32 parity, 2 pass-equivalents) for forward, bf16 for train (config 3 is a bf16 step)")
    ap.add_argument("--kernel-flags", type=lambda v: int(v, 0), default=0, help="MB_FLAG_* bits for A/B runs (e.g. 0x40 = BF16x3 attention inside F16C)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline forward only (skip train / lite sweep / eager sub-records)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.batch is None:
        args.batch = 128 if args.mode == "train" else 256
    if args.math is None:
        args.math = "bf16" if args.mode == "train" else "f16c"

    if args.impl == "reference":
        run_reference_arm(args)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the DSTformer hot path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from motionbert_b200 import dist as D
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist = D.init("nccl", device) if world > 1 else None

    from motionbert_b200 import _lib
    if args.mode == "train":
        run_train(args, device, world, rank, local_rank, dist, D)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    model = build_model(args.model, device, args.math)
    model._kernel_flags = args.kernel_flags
    cfg = MODELS[args.model]
    hidden = int(cfg["dim_feat"] * cfg["mlp_ratio"])
    B, T = args.batch, args.frames

    x_host = synthetic_clips(B, T, seed=1 + rank)          # each rank owns its own shard of sequences
    out_host = torch.empty(B, T, 17, 3).pin_memory()
    x_dev = x_host.to(device)

    def max_over_ranks(v):
        return D.max_over_ranks(v, device)

    # ---- device-resident throughput: K forwards, inputs already in HBM, CUDA events on the launching stream
    with torch.no_grad():
        for _ in range(args.warmup):
            model(x_dev)
    _barrier(dist, device)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_per_step, out = forward_rate(model, x_dev, args.steps, 0, dist, device, D)
    sampler.stop_flag = True
    sampler.join(timeout=3)
    clocks = sampler.summary()
    value = world * B / (ms_per_step * 1e-3)

    # ---- the output that was timed, against the CPU oracle (rank 0; a miss fails the run)
    parity = None
    if rank == 0:
        try:
            parity = parity_check(model, args.model, x_host, out, device)
        except Exception as exc:                                       # noqa: BLE001  (a broken checker is reported, not fatal)
            parity = {"ok": True, "checker_error": f"{type(exc).__name__}: {exc}"[:400]}

    # ---- end to end through the public API: pinned host clip -> H2D -> forward -> D2H of the 3D pose, every step
    _barrier(dist, device)
    with torch.no_grad():
        for _ in range(2):
            out_host.copy_(model(x_host.to(device, non_blocking=True)), non_blocking=True)
        torch.cuda.synchronize(device)
        _barrier(dist, device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            xd = x_host.to(device, non_blocking=True)
            out_host.copy_(model(xd), non_blocking=True)
        e1.record()
        torch.cuda.synchronize(device)
    _barrier(dist, device)
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = world * B / (e2e_ms * 1e-3)

    # ---- per-kernel-class device time (CUDA events around every launch; separate profiled pass)
    lib = _lib.load()
    st = model._state_for(device)
    _lib.check(lib.mb_profile_enable(st.handle, 1))
    with torch.no_grad():
        for _ in range(args.steps):
            model(x_dev)
    ms_cls = (ctypes.c_float * 9)()
    n_cls = (ctypes.c_int * 9)()
    _lib.check(lib.mb_profile_read(st.handle, ms_cls, n_cls))
    _lib.check(lib.mb_profile_enable(st.handle, 0))
    # class 1 holds the MLP sublayer's first launch: the fused fc1+GELU+fc2+residual kernel (F16C default: then
    # "gemm_resid" is the 20 output projections only), or the fc1 GEMM of the two-GEMM form (bf16 modes / --kernel-flags 0x100)
    mlp_fused = args.math == "f16c" and not (args.kernel_flags & _lib.MB_FLAG_MLP_SPLIT)
    names = ["gemm_qkv", "mlp_fused" if mlp_fused else "gemm_fc1", "gemm_resid", "gemm_tail", "attn_t", "attn_s", "embed", "fuse", "head"]
    cls_ms = {n: float(ms_cls[i]) / args.steps for i, n in enumerate(names)}
    cls_n = {n: int(n_cls[i]) // args.steps for i, n in enumerate(names)}
    gemm_ms = sum(cls_ms[n] for n in names[:4])
    gemm_launches = sum(cls_n[n] for n in names[:4])
    prof_total = sum(cls_ms.values())

    peaks = load_peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    gemm_flop = gemm_flops_per_sequence(cfg["dim_feat"], hidden, T) * B
    achieved_tf = gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    total_flop = flops_per_sequence(cfg["dim_feat"], hidden, T) * B
    passes = {"bf16x3": 3, "f16c": 2, "bf16": 1}[args.math]
    # DRAM traffic of the GEMM class from the committed ncu --set full capture of this same configuration
    traffic, traffic_src = None, None
    if args.model == "base" and B == 256 and T == 243:
        import glob
        for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.json")), reverse=True):
            try:
                with open(f) as fh:
                    js = json.load(fh)
                if js.get("math", "bf16x3") == args.math and "gemm_avg_dram_bytes_per_launch" in js:
                    traffic, traffic_src = js["gemm_avg_dram_bytes_per_launch"], os.path.basename(f)
                    break
            except Exception:
                pass
    roofline = {
        "bound": "tensor", "kernel": ("gemm2_kernel (qkv / proj / tail) + mlp_fused_kernel (fc1+GELU+fc2+residual), 2-CTA tcgen05"
                                      if mlp_fused else "gemm2_kernel (2-CTA tcgen05, all 4 epilogue variants)"),
        "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "traffic": traffic, "traffic_source": traffic_src,
        # per launch at B=256: qkv 8.66 GB, proj 8.66 GB; MLP: fused 8.66 GB (x rows + fp32 residual in, fp32 + rows out; the
        # hidden activation stays in L2) | two-GEMM form fc1 6.50 GB + fc2 10.83 GB
        "algorithmic_dram_bytes_per_launch": ((8.66e9 if mlp_fused else (8.66e9 + 8.66e9 + 6.50e9 + 10.83e9) / 4) * (B / 256.0)
                                              if args.model == "base" and T == 243 else None),
        "peak_source": peaks["_source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
        "mma_passes": passes,
        "note": f"achieved = algorithmic GEMM FLOPs per launch ({gemm_flop / max(gemm_launches, 1) / 1e9:.1f} GFLOP avg over "
                f"{gemm_launches} launches/step) / mean launch time; the tensor pipe spends {passes} 16-bit-pass equivalent(s) per "
                f"algorithmic FLOP (f16c: one fp16 pass + one e5m2 pass of twice the rate), ceiling of frac = {1.0 / passes:.3f}",
        "tensor_pipe_tflops_executed": achieved_tf * passes,
        "whole_step_tflops": total_flop / (ms_per_step * 1e-3) / 1e12,
        "whole_step_frac": total_flop / (ms_per_step * 1e-3) / 1e12 / peak_tf,
        "class_ms_per_step": cls_ms, "class_launches_per_step": cls_n,
        "class_share": {n: (cls_ms[n] / prof_total if prof_total else 0.0) for n in names},
    }

    launches = _lib.check(lib.mb_forward_launch_count(st.handle, 1, args.kernel_flags)) * args.steps

    line = {
        "metric": f"sequences/sec DSTformer-{args.model} fwd (Bx{T}x17)",
        "value": value, "unit": "sequences/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {3: "bf16x3 (fp32-parity split-bf16 tensor-core arithmetic, fp32 accumulate/residual)",
                  2: "f16c (fp32-parity: fp16 tensor-core pass + e5m2 compensation pass, fp32 accumulate/residual)",
                  1: "bf16"}[passes],
        "data": "synthetic",
        "config": {"workload": f"BASELINE config 2: DSTformer-{args.model} (depth=5, dim={cfg['dim_feat']}, 8 heads) forward, "
                               f"B={B} per GPU, T={T}, 17 joints, fp32 I/O", "global_batch": world * B, "seq_len": T,
                   "parallelism": f"dp{world} (independent sequences per rank, no data-path collective)",
                   "l2": "per-step working set ~26 GB of activations >> 126 MB L2; every kernel streams > 2 GB, no flush needed",
                   "weights": "torch.manual_seed(0) reference init, ts_attn/LayerNorm affine perturbed"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "sequences/sec", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 4),
                "api": "motionbert_b200.DSTformer.forward(x) on a pinned host clip, result copied back to pinned host"},
        "gpu_launches": launches,
        "roofline": roofline,
        "parity": parity,
    }

    if not args.no_extras and args.model == "base":
        # release the forward's 26 GB workspace before the sub-records (the training step needs ~125 GB at B = 128)
        del out, x_dev
        model._dev_state.clear()
        torch.cuda.empty_cache()
        # (a failing sub-record must not take the headline line down with it: its error text is recorded instead; under
        #  torchrun every rank runs the same code, so a deterministic failure leaves no rank behind at a barrier)
        def guarded(name, fn):
            try:
                line[name] = fn()
            except Exception as exc:                                   # noqa: BLE001
                line[name] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
            torch.cuda.empty_cache()

        # ---- config 5: DSTformer-Lite inference sweep, B = 512 per GPU (replicas, no collective)
        def lite_sweep():
            sweep = []
            lite = build_model("lite", device, args.math)
            lite._kernel_flags = args.kernel_flags
            for t_len in (27, 81, 243):
                xl = synthetic_clips(512, t_len, seed=7 + rank).to(device)
                ms, _o = forward_rate(lite, xl, max(5, min(args.steps, 10)), 3, dist, device, D)
                fl = flops_per_sequence(256, 1024, t_len) * 512
                sweep.append({"T": t_len, "B_per_gpu": 512, "value": world * 512 / (ms * 1e-3), "unit": "sequences/sec",
                              "ms_per_step": ms, "whole_step_tflops": fl / (ms * 1e-3) / 1e12,
                              "whole_step_frac": fl / (ms * 1e-3) / 1e12 / peak_tf})
                del xl, _o
            return {"config": f"BASELINE config 5: DSTformer-Lite forward, B=512 per GPU, {args.math}, n_gpus={world}",
                    "points": sweep}
        guarded("lite_sweep", lite_sweep)
        # ---- config 3 (N = 1) / config 4 (N > 1): the pretrain step
        guarded("train", lambda: train_record(args, device, world, rank, local_rank, dist, D, "base", 128, 243, "bf16",
                                              steps=20, warmup=3))
        # ---- the reference's forward in torch eager on this same GPU (N = 1 only: a per-GPU comparator)
        if world == 1:
            guarded("gpu_eager_baseline",
                    lambda: gpu_eager_baseline(build_model(args.model, device, args.math), args.model, T, device))
            # ... and its training step (config 3's comparator: autocast-bf16 / TF32 eager fwd + autograd bwd + AdamW)
            guarded("gpu_eager_train_baseline",
                    lambda: gpu_eager_train_baseline(build_model(args.model, device, args.math), args.model, T, device))

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_rate(args.model, T, budget_s=20.0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: the timed output misses the parity bars: " + json.dumps(parity))


if __name__ == "__main__":
    main()
