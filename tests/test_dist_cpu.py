"""N>1 path on CPU: world_size-2 gloo processes exercise the sharding + max-over-ranks logic bench.py uses."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from motionbert_b200 import dist as D


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 256, 1024, 1023):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist = D.init("gloo")
    try:
        from oracle import dstformer_oracle as O
        from oracle import dstformer_torch_cpu as OT
        # each rank processes its shard of a global batch of independent sequences (CPU oracle stands in for the GPU)
        cfg = O.LITE
        P = {k: torch.from_numpy(v) for k, v in O.make_params(cfg, 3).items()}
        x = torch.from_numpy(O.make_input(3, 4, cfg.num_joints, 5))
        lo, hi = D.shard_bounds(3, rank, world)
        out, _ = OT.forward(P, x[lo:hi], cfg.depth, cfg.num_heads, cfg.eps)
        slow = D.max_over_ranks(10.0 + rank)
        total = D.sum_over_ranks(float(hi - lo))
        dist.barrier()
        q.put((rank, lo, hi, out.double().sum().item(), slow, total, D.env_world()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_forward_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import dstformer_oracle as O
    from oracle import dstformer_torch_cpu as OT
    cfg = O.LITE
    P = {k: torch.from_numpy(v) for k, v in O.make_params(cfg, 3).items()}
    x = torch.from_numpy(O.make_input(3, 4, cfg.num_joints, 5))
    full, _ = OT.forward(P, x, cfg.depth, cfg.num_heads, cfg.eps)
    assert [(r[1], r[2]) for r in res] == [(0, 2), (2, 3)]
    assert abs(sum(r[3] for r in res) - full.double().sum().item()) < 1e-3      # shards reproduce the whole batch
    assert all(r[4] == 11.0 for r in res)                                        # max over ranks
    assert all(r[5] == 3.0 for r in res)                                         # whole-job unit count
    assert [r[6] for r in res] == [(2, 0, 0), (2, 1, 1)]


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist = D.init("gloo")
    try:
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
        ps[0].grad = torch.full((3, 5), float(rank + 1))
        ps[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
        n = D.allreduce_gradients(ps)                     # ps[2] has no gradient (frozen / unused): skipped
        q.put((rank, n, ps[0].grad.clone(), ps[1].grad.clone(), ps[2].grad))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_is_the_mean_over_ranks():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, g0, g1, g2 in res:
        assert n == 22 and g2 is None
        assert torch.equal(g0, torch.full((3, 5), 1.5))
        assert torch.equal(g1, torch.arange(7, dtype=torch.float32) * 1.5)


def test_gradient_allreduce_is_a_noop_without_a_process_group():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    assert D.allreduce_gradients([p]) == 0 and torch.equal(p.grad, torch.ones(4))
