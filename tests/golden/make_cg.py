"""Record the REAL reference's conjugate gradient (``ConjugateGradientSolver``, pykeops/pykeops/common/operations.py:90-114)
on dense symmetric positive definite systems, with the number of operator applications it made, for
tests/test_frontend.py::test_conjugate_gradient_matches_reference.  Same environment as make_golden.py; nothing compiles.

    PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops python tests/golden/make_cg.py
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from pykeops.common.operations import ConjugateGradientSolver
    g = torch.Generator().manual_seed(424242)
    rec = {}
    for name, n, d, dtype, alpha, eps in (("f64", 60, 2, torch.float64, 0.1, 1e-8), ("f32", 80, 3, torch.float32, 0.5, 1e-5),
                                          ("f32_loose", 50, 1, torch.float32, 1.0, 1e-2), ("zero_rhs", 20, 2, torch.float64, 0.1, 1e-6)):
        x = torch.rand(n, 3, generator=g, dtype=dtype)
        K = torch.exp(-((x[:, None] - x[None]) ** 2).sum(-1) / 0.25) + alpha * torch.eye(n, dtype=dtype)
        b = torch.rand(n, d, generator=g, dtype=dtype) * (0.0 if name == "zero_rhs" else 1.0)
        calls = [0]

        def linop(v):
            calls[0] += 1
            return K @ v

        a = ConjugateGradientSolver("torch", linop, b, eps=eps)
        rec[name + ".K"], rec[name + ".b"], rec[name + ".a"] = K.numpy(), b.numpy(), a.numpy()
        rec[name + ".eps"], rec[name + ".calls"] = np.float64(eps), np.int64(calls[0])
    np.savez_compressed(os.path.join(HERE, "cg.npz"), **rec)
    print({k: int(rec[k]) for k in rec if k.endswith(".calls")})


if __name__ == "__main__":
    main()
