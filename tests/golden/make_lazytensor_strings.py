"""Record the formula strings that the REAL reference front-end (pykeops.torch.LazyTensor, v2.3) builds for a list of
symbolic expressions, so that the test-suite can check that keops_b200's LazyTensor maps every method to the same
operation tree (tests/test_lazytensor_strings.py).  Run where /root/reference is mounted:

    CXX=/usr/bin/g++ KEOPS_CACHE_FOLDER=/tmp/keops_cache \
    PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops python tests/golden/make_lazytensor_strings.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> python expression over the symbolic LazyTensors x (Vi(0,3)), y (Vj(1,3)), b (Vj(2,1)), p (Pm(3,1)), A (Vj(4,9)),
# s (Vi(5,1)); `LT` is the LazyTensor class
EXPRS = {
    "gauss": "(-((x - y) ** 2).sum(-1)).exp() * b",
    "sqdist_methods": "x.sqdist(y) + (x - y).sqnorm2() + (x - y).norm2() + (x | y)",
    "laplace": "(-(x - y).norm2() / p).exp() * b",
    "cauchy": "b / (1 + x.sqdist(y) * p)",
    "unary_a": "(x - y).abs().sum(-1) + ((x * y).relu()).sum(-1) + (x - y).sign().sum(-1) + (x - y).step().sum(-1)",
    "unary_b": "(1 + x.sqdist(y)).log() * (1 + x.sqdist(y)).sqrt() * (1 + x.sqdist(y)).rsqrt() + (1 + x.sqdist(y)).xlogx()",
    "trig": "(x | y).sin() * (x | y).cos() + ((x | y) / 4).asin() + ((x | y) / 4).acos() + (x | y).atan() + (x - y).sinc().sum(-1)",
    "atan2": "(x - y).atan2(x + y).sum(-1)",
    "powers": "((x - y) ** 3).sum(-1) + (1 + x.sqdist(y)) ** 0.5 + (1 + x.sqdist(y)) ** (-1) + (1 + x.sqdist(y)).power(b)",
    "round": "(3 * (x - y)).round(1).sum(-1) + (3 * (x - y)).round().sum(-1)",
    "clamp": "(x - y).clamp(-0.5, 0.5).sum(-1) + (3 * (x - y)).clamp(-1, 1).sum(-1) + (x - y).clamp(b - 1, b).sum(-1)",
    "ifelse_mod": "(x - y).ifelse(x * y, x + y).sum(-1) + (5 * (x - y)).mod(2, -1).sum(-1)",
    "compare": "((x < y) + 2 * (x <= y) + 4 * (x > y) + 8 * (x >= y) + 16 * (x == y) + 32 * (x != y)).sum(-1)",
    "normalize": "(x.normalize() | y) * (x - y).normalize().sum(-1)",
    "weighted": "x.weightedsqdist(y, s) + x.weightedsqdist(y, x) + (x - y).weightedsqnorm(A)",
    "elem_extract": "x.elem(1) * y.extract(1, 2).sum(-1) + x[0] * y[1:3].sum(-1) + y.extract(0, 1)",
    "concat": "x.concat(y).sum(-1) + LT.concatenate((x, y, b)).sum(-1) + LT.cat((x, b), -1).sum(-1)",
    "one_hot": "(3 * b).round().one_hot(6).sum(-1) * x.elem(0)",
    "matvec": "A.matvecmult(x).sum(-1) + x.vecmatmult(A).sum(-1) + x.tensorprod(y).sum(-1)",
    "minmax_vec": "(x - y).max(-1) + (x - y).min(-1) + (x * y).argmax(-1) + (x * y).argmin(-1)",
    "neg_div": "(-x / (1 + y * y)).sum(-1) + (2 / (1 + x.sqdist(y))) - (x.sum(-1) - 3) * 0.5",
    "radd_rsub": "(1 + x).sum(-1) + (1 - x).sum(-1) + (2 * x).sum(-1) + (x * 2.5).sum(-1) + (3 / (1 + x * x)).sum(-1)",
    "grad": "((-x.sqdist(y)).exp() * b).grad(x, s)",
    "square_abs": "(x - y).square().sum(-1) * abs(x - y).sum(-1)",
    "sum_dim2": "(x * y).sum(2) + (x * y).sum(dim=-1)",
    "transpose": "((-x.sqdist(y)).exp() * b).t()",
    "transpose_T": "(x.weightedsqdist(y, s) * b).T",
    "diff": "((-x.sqdist(y)).exp() * b).diff(x, x - y)",
    "divergence": "(x * (-x.sqdist(y)).exp()).divergence(x)",
    "laplacian": "(-x.sqdist(y)).exp().laplacian(x)",
    # (grad_matrix is left out: the reference's method formats the LazyTensor object itself into the string,
    #  lazy_tensor.py:2001-2005, so there is no valid reference string to compare with)
    "bspline_softdtw": "A.bspline(s, 2).sum(-1) + A.bspline(b).sum(-1) + x.softdtw_sqdist(y, p) + A.softdtw_sqdist(x, 0.5)",
    # complex LazyTensors: z (Vj, 2 complex numbers), w (Vi, 2 complex numbers), c (Vi, one complex number)
    "cplx_arith": "z * w + z / w - w * c + (z - c) * b",
    "cplx_mixed": "z * b + x.real2complex().sum(-1) * c + (b * z) - y.extract(0, 2).imag2complex() + z * (p * b).exp1j()",
    "cplx_parts": "z.real * w.imag + abs(z) + z.angle() + (z.conj() * w).real + z.exp().imag + (z * w).sum(-1).real",
    "cplx_python_scalars": "(2j * b) * z + (b * (1 + 2j)) * z + (-2 * 3.14 * 1j * p * b).exp() * z",
    "trace_operator": "(x * s + x / (1 + y.sqnorm2())).trace_operator(x)",
    "tensordot": "A.keops_tensordot(x, (3, 3), (3,), (1,), (0,)).sum(-1) + x.keops_tensordot(y, (3,), (3,), (), (), (1, 0)).sum(-1)",
    "kron": "A.keops_kron(x.tensorprod(y), [3, 3], [3, 3]).sum(-1) + x.keops_kron(y, (3,), (3,)).sum(-1)",
    "factorize": "((-x.sqdist(y)).exp() + (-x.sqdist(y)).exp() * b).factorize((-x.sqdist(y)).exp())",
    "auto_factorize": "((-x.sqdist(y)).exp() + (-x.sqdist(y)).exp() * b).auto_factorize()",
}
REDUCTIONS = {
    "sum1": ("gauss", "sum", dict(dim=1)),
    "sum0": ("gauss", "sum", dict(axis=0)),
    "lse": ("laplace", "logsumexp", dict(dim=1)),
    "lse_w": ("sqdist_methods", "logsumexp", dict(dim=1, weight="b")),
    "ssw": ("sqdist_methods", "sumsoftmaxweight", dict(weight="b", dim=1)),
    "min": ("sqdist_methods", "min", dict(dim=1)),
    "argmin0": ("sqdist_methods", "argmin", dict(dim=0)),
    "max_argmax": ("sqdist_methods", "max_argmax", dict(dim=1)),
    "kmin": ("sqdist_methods", "Kmin", dict(K=4, dim=1)),
    "argkmin": ("sqdist_methods", "argKmin", dict(K=3, dim=0)),
    "kmin_argkmin": ("sqdist_methods", "Kmin_argKmin", dict(K=2, dim=1)),
    "sum_reduction": ("gauss", "sum_reduction", dict(axis=1)),
    "logsumexp_reduction": ("laplace", "logsumexp_reduction", dict(dim=0)),
    "min_argmin_reduction": ("sqdist_methods", "min_argmin_reduction", dict(axis=1)),
    "argKmin_reduction": ("sqdist_methods", "argKmin_reduction", dict(K=6, axis=1)),
    "transposed_sum": ("transpose", "sum", dict(dim=1)),
}


def symbols(LT):
    return dict(x=LT((0, 3, 0)), y=LT((1, 3, 1)), b=LT((2, 1, 1)), p=LT((3, 1, 2)), A=LT((4, 9, 1)), s=LT((5, 1, 0)), LT=LT,
                z=LT((6, 4, 1), is_complex=True), w=LT((7, 4, 0), is_complex=True), c=LT((8, 2, 0), is_complex=True))


def build(LT):
    env = symbols(LT)
    out = {}
    for name, src in EXPRS.items():
        f = eval(src, dict(env))
        out["expr:" + name] = dict(src=src, formula=f.formula.replace("VarSymb", "Var"), ndim=int(f.ndim),
                                   is_complex=bool(getattr(f, "is_complex", False)), shape=list(f.shape))
    for name, (expr, method, kw) in REDUCTIONS.items():
        f = eval(EXPRS[expr], dict(env))
        kw = dict(kw)
        if "weight" in kw:
            kw["weight"] = env[kw["weight"]]
        args = (kw.pop("K"),) if ("K" in kw and not method.endswith("_reduction")) else ()   # *_reduction(**kwargs) take K= by name
        r = getattr(f, method)(*args, call=False, **kw)
        out["red:" + name] = dict(expr=expr, method=method, kwargs={k: (v if not hasattr(v, "formula") else "b") for k, v in kw.items()},
                                  K=(args[0] if args else kw.get("K")), formula=r.formula.replace("VarSymb", "Var"),
                                  formula2=(r.formula2.replace("VarSymb", "Var") if getattr(r, "formula2", None) else None),
                                  reduction_op=r.reduction_op, axis=int(r.axis), opt_arg=getattr(r, "opt_arg", None))
    return out


if __name__ == "__main__":
    from pykeops.torch import LazyTensor
    rec = build(LazyTensor)
    with open(os.path.join(HERE, "lazytensor_formulas.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("recorded %d entries" % len(rec))
