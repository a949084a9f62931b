"""GPU: the pre-scaled / packed variant of Gaussian-type sums in low dimension (``emit._prescale_rewrite``,
``binder._launch_prescaled``): same results as the exact-order kernel and the oracle within the parity bar, taken only for
large problems whose run-time error bound is small, refused (bit-identical fallback) otherwise."""
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
GP = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,2,1),0)"
GRAD = "Grad_WithSavedForward(%s,%%s,Var(3,1,0),Var(4,1,0))" % G


def _run(formula, nargs, args, nx, ny, dev, prescale):
    from keops_b200.binder import get_binder
    os.environ["KB200_PRESCALE"] = "1" if prescale else "0"
    try:
        bd = get_binder(formula, [], nargs, "float32")
        out = bd.genred_pytorch(dev, None, nx, ny, 0, None, *args)
        torch.cuda.synchronize()
        return out, dict(bd.last_launch)
    finally:
        os.environ.pop("KB200_PRESCALE", None)


@pytest.mark.parametrize("case", ["fwd", "grad_x", "grad_y", "grad_b", "param"])
def test_prescaled_kernel_matches_exact_kernel_and_oracle(case, cuda_device):
    from oracle import keops_oracle as ko
    dev = cuda_device
    torch.manual_seed(1)
    M, N = 160000, 150000                                  # 2.4e10 pairs: just above the threshold of the pre-scaled path
    x, y = torch.randn(M, 3, device=dev), torch.randn(N, 3, device=dev)
    if case == "param":
        f, nargs = GP, 4
        args = (x, y, torch.randn(N, 2, device=dev), torch.tensor([0.6], device=dev))
    else:
        b, eta = torch.randn(N, 1, device=dev), torch.randn(M, 1, device=dev)
        fwd, _ = _run(G, 3, (x, y, b), M, N, dev, False)
        if case == "fwd":
            f, nargs, args = G, 3, (x, y, b)
        else:
            var = {"grad_x": "Var(0,3,0)", "grad_y": "Var(1,3,1)", "grad_b": "Var(2,1,1)"}[case]
            f, nargs, args = GRAD % var, 5, (x, y, b, eta, fwd)
    fast, ll = _run(f, nargs, args, M, N, dev, True)
    assert ll.get("engine") == "cuda-core/prescaled", ll
    assert 0 < ll["error_bound"] <= 3e-5
    exact, ll2 = _run(f, nargs, args, M, N, dev, False)
    assert ll2.get("engine") is None
    scale = util.gpu_abs_scale(f, nargs, "float32", dev, M, N, *args).cpu().numpy()
    util.assert_close(fast.cpu().numpy(), exact.cpu().numpy(), "float32", "prescaled vs exact-order kernel, " + case, scale=scale)
    # sampled rows against the numpy oracle (reference operation order)
    tagI = 1 if case in ("grad_y", "grad_b") else 0
    rows = np.arange(0, 40)
    host = [a.cpu().numpy() for a in args]
    from keops_b200.formulas.parser import parse_reduction
    cats = {v.params[0]: v.params[2] for v in parse_reduction(f, []).variables()}
    sub = [a[rows] if cats.get(k) == tagI else a for k, a in enumerate(host)]
    ref = ko.genred(f, [], sub, "float32")
    util.assert_close(fast.cpu().numpy()[rows], ref, "float32", "prescaled vs oracle, " + case, scale=scale[rows])


def test_prescale_guard_refuses_wide_clouds_and_negative_scales(cuda_device):
    """Clouds much wider than the kernel bandwidth would lose digits in the scaled coordinates; a negative scale has no real
    square root: both go to the exact-order kernel, whose result comes back bit for bit."""
    dev = cuda_device
    torch.manual_seed(2)
    M, N = 160000, 150000
    x, y, b = 40.0 * torch.randn(M, 3, device=dev), 40.0 * torch.randn(N, 3, device=dev), torch.randn(N, 1, device=dev)
    o, ll = _run(G, 3, (x, y, b), M, N, dev, True)
    assert ll.get("engine") is None, ll
    o2, _ = _run(G, 3, (x, y, b), M, N, dev, False)
    assert torch.equal(o, o2)
    # shifted (not wider) clouds are fine: the common shift removes the offset
    xs, ys = torch.randn(M, 3, device=dev) + 300.0, torch.randn(N, 3, device=dev) + 300.0
    o3, ll3 = _run(G, 3, (xs, ys, b), M, N, dev, True)
    assert ll3.get("engine") == "cuda-core/prescaled", ll3
    o4, _ = _run(G, 3, (xs, ys, b), M, N, dev, False)
    scale = util.gpu_abs_scale(G, 3, "float32", dev, M, N, xs, ys, b).cpu().numpy()
    util.assert_close(o3.cpu().numpy(), o4.cpu().numpy(), "float32", "prescaled, offset clouds", scale=scale)
    # negative parameter: exp(+|g| d^2) -- refused
    x1, y1 = 0.05 * torch.randn(M, 3, device=dev), 0.05 * torch.randn(N, 3, device=dev)
    args = (x1, y1, torch.randn(N, 2, device=dev), torch.tensor([-0.5], device=dev))
    o5, ll5 = _run(GP, 4, args, M, N, dev, True)
    assert ll5.get("engine") is None, ll5
    o6, _ = _run(GP, 4, args, M, N, dev, False)
    assert torch.equal(o5, o6)


def test_small_problems_keep_the_exact_order_kernel(cuda_device):
    dev = cuda_device
    x, y, b = torch.randn(3000, 3, device=dev), torch.randn(4000, 3, device=dev), torch.randn(4000, 1, device=dev)
    _, ll = _run(G, 3, (x, y, b), 3000, 4000, dev, True)
    assert ll.get("engine") is None
