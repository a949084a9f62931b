"""CPU: host-side logic of the engine -- parser, symbolic autodiff, lowering to the scalar DAG, sign /
FMA rewriting -- evaluated with a numpy interpreter of the DAG (tests/dag_eval.py) against golden
outputs of the real reference, plus metadata (dimout, tagIJ) checks and NVRTC cross-compilation of every
generated kernel for sm_100a."""
import numpy as np
import pytest

from tests import util, dag_eval


@pytest.mark.parametrize("name", util.golden_names())
def test_lowered_formula_matches_reference(name):
    meta, args, ref = util.load_case(name)
    got = dag_eval.evaluate(meta["formula"], meta["aliases"], args, "float64", fast=False, sizes=(meta["nx"], meta["ny"]))
    assert got.shape == ref.shape, (got.shape, ref.shape)
    head = meta["formula"].split("(")[0]
    if "Arg" in head or "KMin" in head:
        # orderings are checked bit-exactly on the GPU and in the oracle; here values / indices in fp64
        same = np.isclose(got, ref, rtol=1e-5, atol=1e-6)
        assert same.mean() > 0.999
    else:
        rtol = 2e-5 if meta["dtype"] == "float32" else util.rtol_for(meta) * 50
        util.assert_close(got, ref, "float64", name, rtol=rtol, **util.golden_scale(meta, args))


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name", ["gauss_grad_x_float32", "lse_grad_x_float32", "kernelsolve_matvec_float32",
                                  "op_div_scalar_float32", "grad_inv_log_x_float64"])
def test_fast_math_regrouping_preserves_values(name, fast):
    meta, args, ref = util.load_case(name)
    got = dag_eval.evaluate(meta["formula"], meta["aliases"], args, "float64", fast=fast)
    util.assert_close(got, ref, "float64", name, rtol=2e-5, **util.golden_scale(meta, args))


def test_metadata_matches_reference():
    from keops_b200.formulas.parser import parse_reduction
    for name in util.golden_names():
        meta, _, _ = util.load_case(name)
        red = parse_reduction(meta["formula"], meta["aliases"])
        assert red.dim == meta["dimout"], name
        assert red.tagI == meta["tagIJ"], name


def test_second_order_gradient_strings():
    """Backward of a backward: the reduction string nests two Grad_WithSavedForward calls
    (pykeops/pykeops/torch/generic/generic_red.py:189); check it against dense numpy calculus."""
    rng = np.random.default_rng(3)
    M, N = 23, 31
    x, y, b = rng.random((M, 3)), rng.random((N, 3)), rng.random((N, 1))
    e1, e2 = rng.standard_normal((M, 1)), rng.standard_normal((N, 1))
    G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
    g1 = "Grad_WithSavedForward(%s,Var(2,1,1),Var(3,1,0),Var(4,1,0))" % G            # d/db -> indexed by j
    g2 = "Grad_WithSavedForward(%s,Var(3,1,0),Var(5,1,1),Var(6,1,1))" % g1           # d/d(eta) of that -> indexed by i
    K = np.exp(-((x[:, None] - y[None]) ** 2).sum(-1))
    fwd = K @ b
    gb = K.T @ e1                       # (N,1)
    got1 = dag_eval.evaluate(g1, [], [x, y, b, e1, fwd])
    np.testing.assert_allclose(got1, gb, rtol=1e-10)
    got2 = dag_eval.evaluate(g2, [], [x, y, b, e1, fwd, e2, gb])
    np.testing.assert_allclose(got2, K @ e2, rtol=1e-10)


def test_every_golden_kernel_cross_compiles_for_sm100a():
    from keops_b200.formulas.parser import parse_reduction
    from keops_b200.codegen.emit import generate
    from keops_b200.runtime import plan
    seen = set()
    for name in util.golden_names():
        meta, _, _ = util.load_case(name)
        key = (meta["formula"], meta["dtype"])
        if key in seen or not name.startswith(("gauss", "argkmin_K10", "lse_softmin", "min_argmin", "kmin_argkmin_f", "kernelsolve")):
            continue
        seen.add(key)
        red = parse_reduction(meta["formula"], meta["aliases"])
        ks = generate(red, meta["dtype"], meta["nargs"], plan.default_config(red, meta["dtype"]))
        if ks.is_zero:
            continue
        cubin = plan.get_cubin(ks, fast=(meta["dtype"] == "float32"))
        assert cubin[:4] == b"\x7fELF", name


def test_parser_errors():
    from keops_b200.formulas.parser import parse_reduction
    from keops_b200.formulas.ir import FormulaError
    for bad in ("Sum_Reduction(Foo(Var(0,3,0)),0)", "Var(0,3,0)+Var(1,3,1)", "Sum_Reduction(Var(0,3,0)+Var(1,2,1),0)",
                "Sum_Reduction(__import__('os').system('true'),0)", "Sum_Reduction(Var(0,3,0),2)"):
        with pytest.raises((FormulaError, ValueError)):
            parse_reduction(bad)


def test_formula_level_grad_equals_reduction_level_grad():
    """``Grad(f, v, e)`` inside a formula (what ``LazyTensor.grad`` prints, lazy_tensor.py:1918-1936) is the same symbolic
    gradient as ``Grad(Sum_Reduction(f), v, e)`` (keopscore/keopscore/formulas/autodiff/Grad.py:4-12)."""
    from keops_b200.formulas.parser import parse_reduction
    f = "Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,1,1)"
    a = parse_reduction("Sum_Reduction(Grad(%s, Var(0,3,0), Var(3,1,0)),0)" % f)
    b = parse_reduction("Grad(Sum_Reduction(%s,0), Var(0,3,0), Var(3,1,0))" % f)
    assert a.formula is b.formula and a.tagI == b.tagI == 0


def test_lazytensor_transposition_bookkeeping():
    """``K.t()`` swaps the i / j roles, sizes and ranges, and relabels its variables (lazy_tensor.py:2529-2601)."""
    import re
    import torch
    from keops_b200.torch import LazyTensor
    x, y = torch.randn(5, 3), torch.randn(7, 3)
    K = (-((LazyTensor(x[:, None, :]) - LazyTensor(y[None, :, :])) ** 2).sum(2)).exp()
    Kt = K.t()
    assert K.shape == (5, 7) and Kt.shape == (7, 5) and K.T.shape == (7, 5)
    cats = lambda f: re.findall(r"Var\((\d+),3,([01])\)", f)
    (xi, ci), (yj, cj) = cats(K.formula)
    (xt, cti), (yt, ctj) = cats(Kt.formula)
    assert (ci, cj, cti, ctj) == ("0", "1", "1", "0")
    assert {xi, yj}.isdisjoint({xt, yt})                       # fresh ids: K and K.t() can share a formula
    assert Kt.t().shape == K.shape
    r = tuple(torch.zeros(1, 2, dtype=torch.int32) + k for k in range(6))
    K.ranges = r
    assert [int(t[0, 0]) for t in K.t().ranges] == [3, 4, 5, 0, 1, 2]


def test_high_dimensional_formula_generates():
    """D = 600: the scalar DAG of a D-term left-to-right sum is D levels deep; the recursive rewriters must cope
    (the reference handles such formulas in chunk mode, keopscore/keopscore/mapreduce/gpu/GpuReduc1D_chunks.py)."""
    from keops_b200.formulas.parser import parse_reduction
    from keops_b200.codegen.emit import generate
    from keops_b200.runtime import plan
    red = parse_reduction("Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,600,0),Var(1,600,1)))*Var(2,1,1),0)")
    ks = generate(red, "float32", 4, plan.default_config(red, "float32"))
    assert "kb200_reduce" in ks.source and ks.stride >= 600


@pytest.mark.parametrize("formula,var", [
    ("Sum_Reduction(Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,2,1),0)", "Var(0,3,0)"),
    ("Sum_Reduction(Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,2,1),0)", "Var(1,3,1)"),
    ("Sum_Reduction(Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,2,1),0)", "Var(2,2,1)"),
    ("Sum_Reduction(Sqrt(IntCst(1)+SqDist(Var(0,3,0),Var(1,3,1)))*Sin(Var(2,2,1))/(IntCst(2)+(Var(0,3,0)|Var(1,3,1))**2),1)", "Var(0,3,0)"),
])
def test_forward_mode_diff_matches_finite_differences(formula, var):
    """``Diff_WithSavedForward`` (forward-mode differential, keopscore/keopscore/formulas/autodiff/Diff.py:14-36): the
    lowered reduction against central finite differences of the forward reduction, in float64."""
    import re
    rng = np.random.default_rng(3)
    M, N = 7, 9
    x, y, b = rng.standard_normal((M, 3)), rng.standard_normal((N, 3)), rng.standard_normal((N, 2))
    args = [x, y, b]
    ind, dim, cat = (int(t) for t in re.findall(r"\d+", var))
    u = rng.standard_normal(args[ind].shape)
    tagI = int(formula.rstrip(")").split(",")[-1])
    fwd = dag_eval.evaluate(formula, [], args, "float64")
    dform = "Diff_WithSavedForward(%s,%s,Var(3,%d,%d),Var(4,%d,%d))" % (formula, var, dim, cat, fwd.shape[1], tagI)
    got = dag_eval.evaluate(dform, [], args + [u, fwd], "float64")
    h = 1e-6
    ap, am = list(args), list(args)
    ap[ind], am[ind] = args[ind] + h * u, args[ind] - h * u
    fd = (dag_eval.evaluate(formula, [], ap, "float64") - dag_eval.evaluate(formula, [], am, "float64")) / (2 * h)
    np.testing.assert_allclose(got, fd, rtol=1e-6, atol=1e-8)


def test_divergence_laplacian_factorize():
    """``Divergence`` / ``Laplacian`` (keopscore/keopscore/formulas/autodiff/Divergence.py, Laplacian.py) against closed forms;
    ``Factorize`` / ``AutoFactorize`` are the identity on a hash-consed DAG."""
    from keops_b200.formulas.parser import parse_formula
    rng = np.random.default_rng(8)
    M, N, D = 6, 5, 3
    x, y = rng.standard_normal((M, D)), rng.standard_normal((N, D))
    g = "Exp(-SqDist(Var(0,3,0),Var(1,3,1)))"
    lap = dag_eval.evaluate("Sum_Reduction(Laplacian(%s,Var(0,3,0)),0)" % g, [], [x, y], "float64")
    d2 = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    np.testing.assert_allclose(lap[:, 0], ((4 * d2 - 2 * D) * np.exp(-d2)).sum(1), rtol=1e-10)
    div = dag_eval.evaluate("Sum_Reduction(Divergence(Var(0,3,0)*Exp(-SqDist(Var(0,3,0),Var(1,3,1))),Var(0,3,0)),0)", [],
                            [x, y], "float64")
    # div_x (x e^{-|x-y|^2}) = (D - 2 x.(x-y)) e^{-|x-y|^2}
    xd = (x[:, None, :] * (x[:, None, :] - y[None, :, :])).sum(-1)
    np.testing.assert_allclose(div[:, 0], ((D - 2 * xd) * np.exp(-d2)).sum(1), rtol=1e-10)
    f = parse_formula("Exp(Var(0,3,0))+Exp(Var(0,3,0))")
    assert parse_formula("Factorize(Exp(Var(0,3,0))+Exp(Var(0,3,0)),Exp(Var(0,3,0)))") is f
    assert parse_formula("AutoFactorize(Exp(Var(0,3,0))+Exp(Var(0,3,0)))") is f


def test_grad_matrix_is_the_jacobian():
    """``GradMatrix(f, v)`` (keopscore/keopscore/formulas/maths/GradMatrix.py:14-28): f = d |d|^2 with d = x - y has the
    Jacobian |d|^2 I + 2 d d^T."""
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((4, 2)), rng.standard_normal((3, 2))
    out = dag_eval.evaluate("Sum_Reduction(GradMatrix((Var(0,2,0)-Var(1,2,1))*SqDist(Var(0,2,0),Var(1,2,1)), Var(0,2,0)),0)",
                            [], [x, y], "float64")
    d = x[:, None, :] - y[None, :, :]
    J = (d ** 2).sum(-1)[..., None, None] * np.eye(2) + 2 * d[..., :, None] * d[..., None, :]
    np.testing.assert_allclose(out, J.sum(1).reshape(4, 4), rtol=1e-12)


@pytest.mark.parametrize("name", ["f64", "f32", "f32_loose", "zero_rhs"])
def test_conjugate_gradient_matches_reference(name):
    """KernelSolve's host loop against recordings of the reference's ``ConjugateGradientSolver``
    (pykeops/pykeops/common/operations.py:90-114; tests/golden/make_cg.py): same iterates bit for bit on a dense
    operator, same number of operator applications (the stopping rule)."""
    import os
    import torch
    from keops_b200.torch.kernel_solve import conjugate_gradient
    rec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cg.npz"))
    K, b, a_ref = (torch.from_numpy(rec["%s.%s" % (name, k)]) for k in ("K", "b", "a"))
    calls = [0]

    def linop(v):
        calls[0] += 1
        return K @ v

    a = conjugate_gradient(linop, b, eps=float(rec[name + ".eps"]))
    assert calls[0] == int(rec[name + ".calls"]) == conjugate_gradient.last_iterations
    assert a.dtype == a_ref.dtype and a.shape == a_ref.shape
    assert torch.equal(a, a_ref)
    capped = conjugate_gradient(linop, b, eps=0.0, max_iter=3) if name != "zero_rhs" else None
    assert capped is None or conjugate_gradient.last_iterations == 3
    # convergence test on the device, host read-back every 4 iterations: same iterates, same count; the wasted operator
    # applications after convergence add exactly zero
    calls[0] = 0
    a_dev = conjugate_gradient(linop, b, eps=float(rec[name + ".eps"]), sync_every=4)
    assert conjugate_gradient.last_iterations == int(rec[name + ".calls"])
    assert int(rec[name + ".calls"]) <= calls[0] <= int(rec[name + ".calls"]) + 3
    assert torch.equal(a_dev, a_ref)


def _constant_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "constants.json")) as f:
        rec = json.load(f)
    return rec["x"], [c for c in rec["cases"] if "out" in c]


@pytest.mark.parametrize("case", _constant_cases()[1], ids=lambda c: c["formula"][:40])
def test_constant_folding_matches_reference(case):
    """Integer / rational constants in float64 against the real reference (tests/golden/make_constants.py): the reference
    folds rationals in Add / Subtract / Mult / Minus / Divide-by-constant only and emits a folded ``RatCst`` float-rounded
    (keopscore/keopscore/formulas/variables/RatCst.py:22-24), so folding in other places -- or not folding where it does --
    shows in the 8th digit.  Both this package's lowering and the oracle must reproduce the reference to the last bits."""
    from oracle import keops_oracle as ko
    x = np.array(_constant_cases()[0]).reshape(-1, 1)
    y = np.zeros((2, 1))
    ref = np.array([float.fromhex(v) for v in case["out"]]).reshape(-1, 1)
    f = "Sum_Reduction(%s+Var(1,1,1),0)" % case["formula"]
    for who, got in (("lowering", dag_eval.evaluate(f, [], [x, y], "float64")), ("oracle", ko.genred(f, [], [x, y], "float64"))):
        assert np.abs(got - ref).max() <= 4e-16 * np.abs(ref).max(), (who, case["formula"], got.ravel().tolist(), ref.ravel().tolist())


@pytest.mark.parametrize("name", ["gauss_sum_float32", "gauss_grad_x_float32", "gauss_grad_y_float32", "gauss_grad_b_float32",
                                  "gauss_sum_axis0_float32", "gauss_randn_200x300_float32", "gauss_vec_b_float32"])
def test_prescaled_rewrite_matches_reference(name):
    """The pre-scaled variant (``emit._prescale_rewrite``: exponent scale folded into shifted + scaled coordinates, one
    FP32 operation fewer per pair) evaluated in float32 -- rounding of the scaled coordinates included -- against the
    reference's outputs: element-wise within 1e-5 of the condition scale; and it really drops the multiply."""
    from keops_b200.codegen.emit import lower_reduction
    from keops_b200.codegen.scalar import IJ, topo_order
    from keops_b200.formulas.parser import parse_reduction
    meta, args, ref = util.load_case(name)
    got = dag_eval.evaluate(meta["formula"], meta["aliases"], args, "float32", fast=True, sizes=(meta["nx"], meta["ny"]), prescale=True)
    assert got is not None, "the Gaussian-type formula %s must match the pre-scale pattern" % name
    util.assert_close(got, ref, "float32", "prescaled " + name, **util.golden_scale(meta, args))
    from keops_b200.binder import _normalise_aliases
    red = parse_reduction(meta["formula"], _normalise_aliases(meta["aliases"]))

    def pair_ops(low):
        roots = [n for _, nodes, _ in low.acc_terms for n in nodes]
        return sum(1 for n in topo_order(roots) if n.cls == IJ and n.op not in ("const", "load"))
    assert pair_ops(lower_reduction(red, True, prescale=True)) == pair_ops(lower_reduction(red, True)) - 1


def test_prescaled_rewrite_with_a_parameter_scale_and_an_offset_cloud():
    """``Exp(-g * SqDist)`` with a run-time parameter (sigma = sqrt(g log2 e), correctly rounded) on clouds far from the
    origin: the common shift keeps the rounding of the scaled coordinates relative to the clouds' extent."""
    from oracle import keops_oracle as ko
    rng = np.random.default_rng(5)
    f = "Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,2,1),0)"
    x = (rng.random((70, 3)) + 500.0).astype(np.float32)
    y = (rng.random((90, 3)) + 500.0).astype(np.float32)
    b = rng.standard_normal((90, 2)).astype(np.float32)
    g = np.array([3.7], np.float32)
    got = dag_eval.evaluate(f, [], [x, y, b, g], "float32", fast=True, prescale=True)
    ref = ko.genred(f, [], [x, y, b, g], "float64")
    util.assert_close(got, ref, "float32", "prescaled, parameter scale, offset clouds", scale=ko.abs_scale(f, [], [x, y, b, g]))


def test_prescale_does_not_apply_to_other_formulas():
    from keops_b200.codegen.emit import lower_reduction
    from keops_b200.formulas.parser import parse_reduction
    for f in ("Sum_Reduction(Exp(-Sqrt(SqDist(Var(0,3,0),Var(1,3,1))))*Var(2,1,1),0)",          # Laplacian: exp of a sqrt
              "Sum_Reduction(Exp(Var(0,3,0)|Var(1,3,1))*Var(2,1,1),0)",                          # exp of a scalar product
              "Sum_Reduction(SqDist(Var(0,3,0),Var(1,3,1))*Var(2,1,1),0)",                       # no exponential
              "Sum_Reduction(Exp(-SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,1,1),0)",               # dimension > 8
              "ArgKMin_Reduction(SqDist(Var(0,3,0),Var(1,3,1)),3,0)"):
        assert lower_reduction(parse_reduction(f), True, prescale=True) is None, f


@pytest.mark.parametrize("da,db,ca,cb,perm", [
    ((4, 7), (7,), (1,), (0,), None),
    ((2, 3, 2, 2, 4), (2, 4, 2, 3, 2, 3), (4, 0, 2), (1, 4, 2), (4, 3, 2, 0, 1)),
    ((2, 3), (2, 2), (), (), (0, 2, 1, 3)),
    ((3, 4), (4, 3), (0, 1), (1, 0), None),
    ((), (4, 3), (), (), (1, 0)),
])
def test_tensordot_layout_and_adjoints_against_numpy(da, db, ca, cb, perm):
    """``TensorDot`` index bookkeeping (keopscore/keopscore/formulas/maths/TensorDot.py:56-139): values against
    ``numpy.tensordot`` + transposition, adjoint parameters against the explicit sums they must reproduce."""
    from keops_b200.formulas import ir
    rng = np.random.default_rng(3)
    lay = ir.TensorDotLayout(da, db, ca, cb, perm)

    def ev(lay, a, b):
        out = np.zeros(lay.dim)
        for o, prods in lay.terms():
            out[o] = sum(a[i] * b[j] for i, j in prods)
        return out

    a, b = rng.standard_normal(int(np.prod(da))), rng.standard_normal(int(np.prod(db)))
    nat = np.tensordot(a.reshape(da), b.reshape(db), axes=(list(ca), list(cb)))
    ref = nat.transpose(np.argsort(lay.permute)) if nat.ndim else nat
    assert list(ref.shape) == lay.out_shape
    np.testing.assert_allclose(ev(lay, a, b), ref.ravel(), rtol=1e-13, atol=1e-13)
    g = rng.standard_normal(lay.dim)
    ga, gb = np.zeros_like(a), np.zeros_like(b)
    for o, prods in lay.terms():
        for i, j in prods:
            ga[i] += g[o] * b[j]
            gb[j] += g[o] * a[i]
    np.testing.assert_allclose(ev(ir.TensorDotLayout(*lay.adjoint(0)), g, b), ga, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(ev(ir.TensorDotLayout(*lay.adjoint(1)), g, a), gb, rtol=1e-13, atol=1e-13)


def test_tensordot_parsing_and_errors():
    from keops_b200.formulas import ir
    from keops_b200.formulas.parser import parse_formula, FormulaError
    al = ["x=Vi(0,6)", "y=Vj(1,4)"]
    f = parse_formula("TensorDot(x, y, [2, 3], [2, 2], [], [], [0, 2, 1, 3])", al)
    assert f is parse_formula("Kron(x, y, [2,3], [2,2])", al) and f.dim == 24
    assert parse_formula(repr(f)) is f                               # the printed form parses back to the same node
    assert parse_formula("TensorDot(x, y, (2, 3), (2, 2), (0,), (1,))", al).dim == 6          # tuples as well as lists
    assert parse_formula("TensorDot(x, Zero(4), [2,3], [2,2], [], [])", al).is_zero
    for bad in ("TensorDot(x, y, [2, 3], [2, 2], [1], [0])",           # contracted sizes differ (3 vs 2)
                "TensorDot(x, y, [2, 2], [2, 2], [], [])",             # shape does not match the operand dimension
                "TensorDot(x, y, [2, 3], [2, 2], [], [], [0, 1, 2, 2])",   # not a permutation
                "TensorDot(x, y, [2, 3], [2, 2], [0], [0, 1])",        # unpaired axes
                "Kron(x, y, [2, 3], [4])",                             # different numbers of axes
                "TensorDot(x, y, [2, 3.5], [2, 2], [], [])"):
        with pytest.raises(FormulaError):
            parse_formula(bad, al)


def test_bspline_and_softdtw_rules():
    """B-splines: partition of unity inside the knot range, argument checks, no derivative (maths/BSpline.py:24-41,97-103).
    Soft-DTW: small-gamma limit = hard DTW by dynamic programming; no derivative with respect to gamma
    (maths/SoftDTW_SqDist.py:95-101)."""
    from keops_b200.formulas import autodiff, ir
    from keops_b200.formulas.parser import parse_formula, FormulaError
    rng = np.random.default_rng(11)
    t = np.sort(rng.random(9))
    x = (t[3] + (t[5] - t[3]) * rng.random((40, 1)))
    for order in (0, 1, 2, 3):
        got = dag_eval.evaluate("Sum_Reduction(BSpline(Var(0,9,2),Var(1,1,0),%d),0)" % order, [], [t, x], "float64", sizes=(40, 1))
        assert got.shape == (40, 9 - order - 1)
        np.testing.assert_allclose(got.sum(1), 1.0, rtol=1e-12)
        assert (got >= 0).all()
    al = ["t=Pm(0,5)", "x=Vi(1,1)", "v=Vi(2,3)"]
    for bad in ("BSpline(t,x,-1)", "BSpline(t,v,1)", "BSpline(t,x,4)"):
        with pytest.raises(FormulaError):
            parse_formula(bad, al)
    with pytest.raises(NotImplementedError):
        autodiff.grad(parse_formula("BSpline(t,x,2)", al), ir.Var(1, 1, 0), ir.Var(3, 2, 0))

    a, b = rng.standard_normal((6, 5)), rng.standard_normal((7, 4))
    got = dag_eval.evaluate("Min_Reduction(SoftDTW_SqDist(Var(0,5,0),Var(1,4,1),Var(2,1,2)),0)", [], [a, b, np.array([1e-3])], "float64")
    for i in range(6):
        best = np.inf
        for j in range(7):
            r = np.full((6, 5), np.inf)
            r[0, 0] = 0.0
            for p in range(1, 6):
                for q in range(1, 5):
                    r[p, q] = (a[i, p - 1] - b[j, q - 1]) ** 2 + min(r[p - 1, q - 1], r[p - 1, q], r[p, q - 1])
            best = min(best, r[5, 4])
        assert abs(got[i, 0] - best) < 5e-3 * (1 + best)
    f = parse_formula("SoftDTW_SqDist(Var(0,5,0),Var(1,4,1),Var(2,1,2))")
    with pytest.raises(FormulaError):
        autodiff.grad(f, ir.Var(2, 1, 2), ir.Var(3, 1, 0))
    with pytest.raises(FormulaError):
        autodiff.grad(autodiff.grad(f, ir.Var(0, 5, 0), ir.Var(3, 1, 0)), ir.Var(1, 4, 1), ir.Var(4, 5, 0))


def test_product_regrouping_terminates_with_one_row_side_and_two_streamed_factors():
    """(a_i * o_j) * w_j: the streamed factors meet (o_j * w_j leaves the pair loop); moving w_j onto a_i gains nothing and
    used to ping-pong with the mirror-image move without end -- found by the reference's own per-operation harness
    (keopscore/keopscore/test/test_op.py) on d/dv of ``Divide(u_i, v_j)`` and on the bounds of ``Clamp``."""
    from keops_b200.codegen.scalar import Graph, I, J, IJ
    g = Graph(fast=True)
    a, w, o = g.load(0, 0, 0), g.load(1, 0, 1), g.load(2, 0, 1)
    r = g.mul(g.mul(a, o), w)
    assert r.n.cls == IJ and sorted(x.cls for x in r.n.args) == [I, J]
    g = Graph(fast=True)
    a, w, o = g.load(0, 0, 1), g.load(1, 0, 0), g.load(2, 0, 0)
    r = g.mul(g.mul(a, o), w)
    assert r.n.cls == IJ and sorted(x.cls for x in r.n.args) == [I, J]
    x, y, e = np.random.default_rng(0).random((3, 7, 3)) + 0.5
    got = dag_eval.evaluate("Grad_WithSavedForward(Sum_Reduction(Divide(Var(0,3,0),Var(1,3,1)),0),Var(1,3,1),Var(2,3,0),Var(3,3,0))",
                            [], [x, y, e, e], "float64", fast=True)
    np.testing.assert_allclose(got, -(e[:, None, :] * x[:, None, :] / y[None, :, :] ** 2).sum(0), rtol=1e-12)


def test_linear_operators_rules():
    """``TraceOperator`` / ``AdjointOperator`` (keopscore/keopscore/formulas/LinearOperators/): linearity is checked by the
    reference's structural rules, the placeholder variable disappears, and the results agree with the explicit matrix."""
    from keops_b200.formulas import autodiff, ir
    from keops_b200.formulas.parser import parse_formula, FormulaError
    al = ["A=Vj(0,9)", "x=Vi(1,3)", "u=Vi(2,3)"]
    t = parse_formula("TraceOperator(MatVecMult(A,u),u)", al)               # (the reference has no rule for MatVecMult here)
    assert [v.params[0] for v in t.variables()] == [0]
    rng = np.random.default_rng(5)
    A, x = rng.standard_normal((6, 9)), rng.standard_normal((4, 3))
    got = dag_eval.evaluate("Sum_Reduction(TraceOperator(MatVecMult(A,u),u)+Sum(AdjointOperator(MatVecMult(A,u),u,x)),1)", al, [A, x],
                            "float64")
    M = A.reshape(6, 3, 3)
    # Sum over i (4 rows) of tr(M_j) + sum_k (M_j^T x_i)_k, one value per j
    want = np.trace(M, axis1=1, axis2=2)[:, None] * 4 + np.einsum("jkl,ik->j", M, x)[:, None]
    np.testing.assert_allclose(got, want, rtol=1e-12)
    for bad in ("TraceOperator(u*u,u)", "TraceOperator(Exp(u),u)", "TraceOperator(u+x,u)", "AdjointOperator(Square(u),u,x)"):
        with pytest.raises(FormulaError, match="not linear"):
            parse_formula(bad, al)
    with pytest.raises(FormulaError):
        parse_formula("TraceOperator(Concat(u,u),u)", al)                    # output dimension neither 1 nor dim(u)
    with pytest.raises(FormulaError):
        parse_formula("AdjointOperator(u*x,u,A)", al)                        # argument of the wrong size
    assert autodiff.is_linear(parse_formula("u/Sum(x)-SumT(Sum(u*x),3)", al), ir.Var(2, 3, 0))
