"""Two-GPU data-parallel training of the drop-in module: one process per GPU over
NCCL; gradients exchanged (a) by motionbert_b200.dist.allreduce_gradients (one flat bucket), (b) by torch's own
DistributedDataParallel wrapper around the unchanged module and (c) by the module's own per-phase exchange overlapped
with the backward (`enable_gradient_allreduce`).  Both must equal the single-process gradient of the whole
batch (up to the bf16 arithmetic of the native backward).  Skipped on a single-GPU box."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_backward import _module
    from motionbert_b200 import dist as D
    from oracle import dstformer_oracle as O
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist = D.init("nccl", dev)
    try:
        m = _module(dev, 256, 1, 8, 2, seed=5)
        x = torch.from_numpy(O.make_input(4, 10, 17, 3)).to(dev)
        w = torch.randn(4, 10, 17, 3, generator=torch.Generator().manual_seed(2)).to(dev)
        lo, hi = D.shard_bounds(4, rank, world)
        # (a) explicit flat-bucket all-reduce
        (m(x[lo:hi]) * w[lo:hi]).sum().backward()
        n = D.allreduce_gradients(list(m.parameters()))
        ga = {k: p.grad.clone() for k, p in m.named_parameters()}
        # (b) torch DDP around the same module
        m.zero_grad(set_to_none=True)
        ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[rank])
        (ddp(x[lo:hi]) * w[lo:hi]).sum().backward()
        gb = {k: p.grad.clone() for k, p in m.named_parameters()}
        # (c) the module's own exchange: per-phase all-reduce on a side stream, overlapped with the backward
        m.zero_grad(set_to_none=True)
        m.enable_gradient_allreduce()
        (m(x[lo:hi]) * w[lo:hi]).sum().backward()
        gc = {k: p.grad.clone() for k, p in m.named_parameters()}
        m.enable_gradient_allreduce(enabled=False)
        # single-process reference on rank 0: whole batch, gradient / world
        full = None
        if rank == 0:
            m.zero_grad(set_to_none=True)
            (m(x) * w).sum().backward()
            full = {k: (p.grad / world).cpu().numpy() for k, p in m.named_parameters()}
        torch.cuda.synchronize(dev)
        dist.barrier()
        # numpy payloads: pickled by value (torch tensors travel as shared-memory handles that die with this process)
        q.put((rank, n, {k: v.cpu().numpy() for k, v in ga.items()}, {k: v.cpu().numpy() for k, v in gb.items()}, full,
               {k: v.cpu().numpy() for k, v in gc.items()}))
    finally:
        dist.destroy_process_group()


def test_two_gpu_gradient_exchange_matches_single_process():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    import numpy as np
    full = res[0][4]
    n_param = sum(v.size for v in full.values())
    assert res[0][1] == n_param and res[1][1] == n_param
    for k, ref in full.items():
        den = float(np.linalg.norm(ref))
        if den == 0:
            continue
        for rank in range(2):
            for which, g in (("flat all-reduce", res[rank][2][k]), ("DDP", res[rank][3][k]), ("overlapped", res[rank][5][k])):
                assert float(np.linalg.norm(g - ref)) / den < 3e-2, f"{which} rank {rank} {k}"
        # both ranks hold the same averaged gradient
        assert np.array_equal(res[0][2][k], res[1][2][k]) and np.array_equal(res[0][5][k], res[1][5][k])
