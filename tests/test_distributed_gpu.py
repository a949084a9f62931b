"""GPU (>= 2 devices): output-index sharding over NCCL with the real CUDA Genred on every rank."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from keops_b200.distributed import ShardedGenred, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    g = torch.Generator().manual_seed(0)
    M, N = 5003, 7001
    x, y, b = torch.rand(M, 3, generator=g), torch.rand(N, 3, generator=g), torch.rand(N, 2, generator=g)
    i0, i1 = shard_bounds(M, world, rank)
    j0, j1 = shard_bounds(N, world, rank)
    sg = ShardedGenred("Exp(-SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)"], "Sum", axis=1)
    out = sg(x[i0:i1].to(dev), y[j0:j1].to(dev), b[j0:j1].to(dev), n_streamed=N, mode="all_gather")
    sk = ShardedGenred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], "ArgKMin", axis=1, opt_arg=5)
    idx = sk(x[i0:i1].to(dev), y.to(dev) if rank == 0 else None, n_streamed=N, mode="broadcast", src=0)
    ref = torch.exp(-((x[i0:i1, None] - y[None]) ** 2).sum(-1)) @ b
    refi = ((x[i0:i1, None] - y[None]) ** 2).sum(-1).topk(5, dim=1, largest=False).indices
    ok = torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-5) and (idx.cpu() != refi).float().mean().item() < 1e-3
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_sharded_reduction(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
