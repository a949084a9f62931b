"""GPU (>= 2 devices): output-index sharding over NCCL with the real CUDA engine on every rank -- forward reductions,
gradients sharded by the differentiated variable's own index (d/dx by i, d/dy by j), LogSumExp, K-min on the tensor-core
path and the row-sharded conjugate gradient, each compared with the single-GPU result of the same engine (bit-identical
where the per-row arithmetic is the same) and with dense torch."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import traceback
    try:
        _worker_body(rank, world, port, q)
    except Exception:                                   # report instead of leaving the parent waiting on the queue
        q.put((rank, ["EXCEPTION on rank %d:\n%s" % (rank, traceback.format_exc()[-3000:])]))


def _worker_body(rank, world, port, q):
    import torch.distributed as dist
    from keops_b200.distributed import ShardedGenred, ShardedKernelSolve, shard_bounds, all_gather_rows
    from keops_b200.torch import Genred, KernelSolve
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    fails = []

    def check(name, cond):
        if not cond:
            fails.append(name)

    g = torch.Generator().manual_seed(0)
    M, N = 5003, 7001
    x, y, b = torch.rand(M, 3, generator=g), torch.rand(N, 3, generator=g), torch.rand(N, 2, generator=g)
    eta = torch.randn(M, 2, generator=g)
    i0, i1 = shard_bounds(M, world, rank)
    j0, j1 = shard_bounds(N, world, rank)
    xd, yd, bd, ed = x.to(dev), y.to(dev), b.to(dev), eta.to(dev)
    al = ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)"]
    # ---- forward Sum: all_gather of the streamed slabs; ArgKMin with a broadcast database
    sg = ShardedGenred("Exp(-SqDist(x,y))*b", al, "Sum", axis=1)
    out = sg(xd[i0:i1], yd[j0:j1], bd[j0:j1], n_streamed=N, mode="all_gather")
    single = Genred("Exp(-SqDist(x,y))*b", al, reduction_op="Sum", axis=1)(xd, yd, bd)
    def near(a, b_):         # same per-pair arithmetic; only the column-split association of the fp32 sums may differ
        return a.shape == b_.shape and bool(((a - b_).abs() <= 2e-6 * b_.abs() + 1e-6 * b_.abs().max()).all())

    check("sum == single-GPU rows", near(out, single[i0:i1]))
    # dense reference in float64 (a float32 matmul on the host carries its own summation error of the same order as the
    # bar); every term is positive, so rtol applies to the sums themselves
    ref = (torch.exp(-((x[i0:i1, None].double() - y[None].double()) ** 2).sum(-1)) @ b.double())
    check("sum vs dense", torch.allclose(out.cpu().double(), ref, rtol=1e-5, atol=1e-6))
    sk = ShardedGenred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], "ArgKMin", axis=1, opt_arg=5)
    idx = sk(xd[i0:i1], yd if rank == 0 else None, n_streamed=N, mode="broadcast", src=0)
    refi = ((x[i0:i1, None] - y[None]) ** 2).sum(-1).topk(5, dim=1, largest=False).indices
    check("argkmin vs dense", (idx.cpu() != refi).float().mean().item() < 1e-3)
    # ---- gradients, each sharded by its own variable's index, against autograd through the single-GPU Genred
    xg, yg, bg = xd.clone().requires_grad_(True), yd.clone().requires_grad_(True), bd.clone().requires_grad_(True)
    full = Genred("Exp(-SqDist(x,y))*b", al, reduction_op="Sum", axis=1)(xg, yg, bg)
    gx, gy, gb = torch.autograd.grad(full, [xg, yg, bg], ed)
    fwd = full.detach()
    loc = (xd[i0:i1], yd[j0:j1], bd[j0:j1], ed[i0:i1], fwd[i0:i1])
    sx = sg.gradient("x")(*loc, n_i=M, n_j=N)
    sy = sg.gradient("y")(*loc, n_i=M, n_j=N)
    sb = sg.gradient("b")(*loc, n_i=M, n_j=N)
    check("grad_x slab (sharded by i)", sx.shape == (i1 - i0, 3) and near(sx, gx[i0:i1]))
    check("grad_y slab (sharded by j)", sy.shape == (j1 - j0, 3) and near(sy, gy[j0:j1]))
    check("grad_b slab (sharded by j)", sb.shape == (j1 - j0, 2) and near(sb, gb[j0:j1]))
    # ---- LogSumExp soft-min with a parameter
    sl = ShardedGenred("f - SqDist(x,y)/IntCst(2)/e", ["f=Vj(1)", "x=Vi(3)", "y=Vj(3)", "e=Pm(1)"], "LogSumExp", axis=1)
    f = torch.rand(N, 1, generator=g).to(dev)
    eps = torch.tensor([0.05], device=dev)
    lse = sl(f[j0:j1], xd[i0:i1], yd[j0:j1], eps, n_streamed=N)
    refl = torch.logsumexp(f.view(1, -1).double() - ((xd[i0:i1, None].double() - yd[None].double()) ** 2).sum(-1) / 2 / 0.05, dim=1)
    check("logsumexp vs dense", torch.allclose(lse.view(-1).double(), refl, rtol=1e-5, atol=1e-5))
    # ---- K-min in D = 64 on the tensor-core filter path, sharded queries
    M2, N2 = 4096, 6000
    q64, d64 = torch.randn(M2, 64, generator=g).to(dev), torch.randn(N2, 64, generator=g).to(dev)
    a0, a1 = shard_bounds(M2, world, rank)
    c0, c1 = shard_bounds(N2, world, rank)
    s64 = ShardedGenred("SqDist(x,y)", ["x=Vi(64)", "y=Vj(64)"], "ArgKMin", axis=1, opt_arg=10)
    i64 = s64(q64[a0:a1].contiguous(), d64[c0:c1].contiguous(), n_streamed=N2)
    one = Genred("SqDist(x,y)", ["x=Vi(64)", "y=Vj(64)"], reduction_op="ArgKMin", axis=1, opt_arg=10)(q64, d64)
    check("argkmin D=64 == single-GPU rows", torch.equal(i64, one[a0:a1]))
    # ---- row-sharded conjugate gradient against the single-GPU KernelSolve
    n = 3000
    xs = torch.rand(n, 3, generator=g).to(dev)
    rhs = torch.randn(n, 1, generator=g).to(dev)
    gam = torch.tensor([30.0], device=dev)
    k0, k1 = shard_bounds(n, world, rank)
    ka = ["x=Vi(3)", "y=Vj(3)", "b=Vj(1)", "g=Pm(1)"]
    sks = ShardedKernelSolve("Exp(-g*SqDist(x,y))*b", ka, "b", axis=1)
    sol = sks(xs[k0:k1], xs[k0:k1], rhs[k0:k1], gam, n_total=n, alpha=0.5, eps=1e-6)
    ref_sol = KernelSolve("Exp(-g*SqDist(x,y))*b", ka, "b", axis=1)(xs, xs, rhs, gam, alpha=0.5, eps=1e-6)
    check("kernel solve", torch.allclose(sol, ref_sol[k0:k1], rtol=1e-4, atol=1e-5 * ref_sol.abs().max().item()))
    check("gathered solution", all_gather_rows(sol, n).shape == (n, 1))
    q.put((rank, fails))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_sharded_reduction(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import queue
    import time
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
    # a rank that raises reports through the queue and exits; its peer then blocks in the next collective -- so collect
    # what arrives, stop waiting as soon as one rank reported a failure, and never leave a worker behind
    res, deadline = [], time.monotonic() + 300
    while len(res) < 2 and time.monotonic() < deadline:
        try:
            res.append(q.get(timeout=5))
        except queue.Empty:
            continue
        if res[-1][1]:
            break
    for p in procs:
        p.join(timeout=10 if len(res) == 2 else 0.1)
        if p.is_alive():
            p.kill()
    assert len(res) == 2 and all(not fails for _, fails in res), res
