"""CPU, world_size = 2, gloo: the output-index sharding and replication logic of
keops_b200.distributed (the compute step is injected: the CPU oracle stands in for the CUDA Genred,
which is exercised on GPUs by tests/test_distributed_gpu.py and bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from keops_b200.distributed import ShardedGenred, all_gather_rows, shard_bounds

FORMULA, ALIASES = "Exp(-SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)"]
RED = "Sum_Reduction(Exp(-SqDist(x,y))*b,0)"
KNN_F, KNN_A = "SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"]


def oracle_compute(red_string, aliases):
    from oracle import keops_oracle as ko

    def run(*args):
        out = ko.genred(red_string, aliases, [a.numpy() for a in args], "float32")
        return torch.from_numpy(out)
    return run


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, N, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x, y, b = torch.rand(M, 3, generator=g), torch.rand(N, 3, generator=g), torch.rand(N, 2, generator=g)
    i0, i1 = shard_bounds(M, world, rank)
    j0, j1 = shard_bounds(N, world, rank)
    sg = ShardedGenred(FORMULA, ALIASES, "Sum", axis=1, compute=oracle_compute(RED, ALIASES))
    if mode == "all_gather":
        out = sg(x[i0:i1], y[j0:j1], b[j0:j1], n_streamed=N, mode="all_gather")
    elif mode == "broadcast":
        yy, bb = (y, b) if rank == 0 else (None, None)
        out = sg(x[i0:i1], yy, bb, n_streamed=N, mode="broadcast", src=0)
    else:
        out = sg(x[i0:i1], y, b, mode="replicated")
    # arg-type reduction: indices must be global without any fix-up
    sk = ShardedGenred(KNN_F, KNN_A, "ArgKMin", axis=1, opt_arg=3,
                       compute=oracle_compute("ArgKMin_Reduction(SqDist(x,y),3,0)", KNN_A))
    idx = sk(x[i0:i1], y[j0:j1], n_streamed=N, mode="all_gather")
    gathered = all_gather_rows(out, M)
    q.put((rank, i0, i1, out.numpy(), idx.numpy(), gathered.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["all_gather", "broadcast", "replicated"])
@pytest.mark.parametrize("M,N", [(37, 53), (40, 64)])
def test_sharded_reduction_matches_single_process(mode, M, N):
    from oracle import keops_oracle as ko
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, N, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    x, y, b = torch.rand(M, 3, generator=g), torch.rand(N, 3, generator=g), torch.rand(N, 2, generator=g)
    full = ko.genred(RED, ALIASES, [x.numpy(), y.numpy(), b.numpy()], "float32")
    knn = ko.genred("ArgKMin_Reduction(SqDist(x,y),3,0)", KNN_A, [x.numpy(), y.numpy()], "float32")
    for rank, i0, i1, out, idx, gathered in res:
        assert np.array_equal(out, full[i0:i1])          # same oracle arithmetic on the same rows: bit-exact
        assert np.array_equal(idx, knn[i0:i1])
        assert np.array_equal(gathered, full)


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------------------------------------
# row-sharded conjugate gradient (KernelSolve over 2 ranks)
# ---------------------------------------------------------------------------------------------
KS_F, KS_A = "Exp(-g*SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)", "g=Pm(1)"]
KS_RED = "Sum_Reduction(Exp(-g*SqDist(x,y))*b,0)"


def _ks_worker(rank, world, port, n, q):
    from keops_b200.distributed import ShardedKernelSolve
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, 3, generator=g, dtype=torch.float64)
    rhs = torch.randn(n, 2, generator=g, dtype=torch.float64)
    gam = torch.tensor([4.0], dtype=torch.float64)
    i0, i1 = shard_bounds(n, world, rank)

    def compute(*args):
        from oracle import keops_oracle as ko
        return torch.from_numpy(ko.genred(KS_RED, KS_A, [a.numpy() for a in args], "float64"))

    ks = ShardedKernelSolve(KS_F, KS_A, "b", axis=1, compute=compute)
    a = ks(x[i0:i1], x[i0:i1], rhs[i0:i1], gam, n_total=n, alpha=0.5, eps=1e-10)
    q.put((rank, i0, i1, a.numpy(), ks.last_iterations))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_kernel_solve_matches_dense_solve():
    world, port, n = 2, _free_port(), 61
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ks_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, 3, generator=g, dtype=torch.float64)
    rhs = torch.randn(n, 2, generator=g, dtype=torch.float64)
    K = torch.exp(-4.0 * ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)) + 0.5 * torch.eye(n, dtype=torch.float64)
    sol = torch.linalg.solve(K, rhs).numpy()
    iters = {r[4] for r in res}
    assert len(iters) == 1 and 0 < iters.pop() <= 2 * n          # every rank stops at the same iteration
    for rank, i0, i1, a, _ in res:
        assert np.abs(a - sol[i0:i1]).max() < 1e-7 * np.abs(sol).max()


# ---------------------------------------------------------------------------------------------
# gradients sharded by the differentiated variable's own index (ShardedGenred.gradient)
# ---------------------------------------------------------------------------------------------
GP_F, GP_A = "Exp(-g*SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(1)", "g=Pm(1)"]
GP_RED = "Sum_Reduction(Exp(-g*SqDist(x,y))*b,0)"


def _grad_inputs(M, N):
    g = torch.Generator().manual_seed(3)
    x, y, b = torch.rand(M, 3, generator=g), torch.rand(N, 3, generator=g), torch.randn(N, 1, generator=g)
    gam, eta = torch.tensor([1.7]), torch.randn(M, 1, generator=g)
    return x, y, b, gam, eta


def _grad_worker(rank, world, port, M, N, q):
    from oracle import keops_oracle as ko
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, y, b, gam, eta = _grad_inputs(M, N)
    fwd = torch.from_numpy(ko.genred(GP_RED, GP_A, [t.numpy() for t in (x, y, b, gam)], "float32"))
    i0, i1 = shard_bounds(M, world, rank)
    j0, j1 = shard_bounds(N, world, rank)
    sg = ShardedGenred(GP_F, GP_A, "Sum", axis=1, compute=oracle_compute(GP_RED, GP_A))
    out = {}
    for var in ("x", "y", "b", "g"):
        gr = sg.gradient(var, compute=lambda *a, _v=var: None)        # placeholder, replaced below (needs gr.red_formula)
        gr.compute = oracle_compute(gr.red_formula, gr.aliases)
        # every array is sharded by its own index; y was replicated beforehand in the 'y' run to exercise per-argument modes
        if var == "x":
            res = gr(x[i0:i1], y, b[j0:j1], gam, eta[i0:i1], fwd[i0:i1], n_i=M, n_j=N, mode={1: "replicated"})
        else:
            res = gr(x[i0:i1], y[j0:j1], b[j0:j1], gam, eta[i0:i1], fwd[i0:i1], n_i=M, n_j=N)
        out[var] = res.numpy()
    q.put((rank, (i0, i1), (j0, j1), out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_match_single_process():
    """d/dx stays sharded by i, d/dy and d/db by j (the i-side arrays, eta and the saved forward are all-gathered), d/dg is the
    all-reduced sum: all equal to the single-process ``Grad_WithSavedForward`` reductions the reference's backward launches."""
    from oracle import keops_oracle as ko
    world, port, M, N = 2, _free_port(), 37, 53
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, M, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y, b, gam, eta = _grad_inputs(M, N)
    host = [t.numpy() for t in (x, y, b, gam)]
    fwd = ko.genred(GP_RED, GP_A, host, "float32")
    full = {}
    for var, vstr in (("x", "Var(0,3,0)"), ("y", "Var(1,3,1)"), ("b", "Var(2,1,1)"), ("g", "Var(3,1,2)")):
        f = "Grad_WithSavedForward(%s,%s,Var(4,1,0),Var(5,1,0))" % (GP_RED, vstr)
        full[var] = ko.genred(f, GP_A + ["Var(4,1,0)", "Var(5,1,0)"], host + [eta.numpy(), fwd], "float32")
    for rank, (i0, i1), (j0, j1), out in res:
        assert np.array_equal(out["x"], full["x"][i0:i1])
        assert np.array_equal(out["y"], full["y"][j0:j1])
        assert np.array_equal(out["b"], full["b"][j0:j1])
        np.testing.assert_allclose(out["g"], full["g"].sum(0), rtol=1e-5)
