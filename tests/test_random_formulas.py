"""CPU: randomized differential test of the formula pipeline.  Seeded random compositions of the reference's operations
(elementwise maths, vector reductions, scalar products, constants, parameters) are pushed through this package's parser,
autodiff and lowering (evaluated with tests/dag_eval.py) and through the oracle (oracle/keops_oracle.py, pinned against
the real reference on 193 fixtures): forward values for Sum / LogSumExp / Min reductions and ``Grad_WithSavedForward``
gradients with respect to an i-variable, two j-variables and a parameter must agree in float64, through the plain and the
``use_fast_math`` lowering.  The 193 fixtures cover each
operation on its own; this covers what only shows up in combinations (common-subexpression sharing, sign pull-out, constant
folding, broadcasting of dimension-1 operands, gradients flowing through several uses of one variable)."""
import random

import numpy as np
import pytest

from oracle import keops_oracle as ko
from tests import dag_eval

X, Y, B, P = "Var(0,3,0)", "Var(1,3,1)", "Var(2,1,1)", "Var(3,1,2)"


def vec(rng, depth):
    """a random dimension-3 expression"""
    if depth <= 0 or rng.random() < 0.25:
        return rng.choice([X, Y, "(%s-%s)" % (X, Y), "(%s+%s)" % (X, Y)])
    k = rng.randrange(11)
    a = vec(rng, depth - 1)
    if k == 0:
        return "(%s*%s)" % (a, vec(rng, depth - 1))
    if k == 1:
        return "(%s-%s)" % (a, vec(rng, depth - 1))
    if k == 2:
        return "(%s*%s)" % (sca(rng, depth - 1), a)
    if k == 3:
        return "Square(%s)" % a
    if k == 4:
        return "Sin(%s)" % a
    if k == 5:
        return "Cos(%s)" % a
    if k == 6:
        return "Exp(Minus(Square(%s)))" % a
    if k == 7:
        return "(%s/(IntCst(1)+Square(%s)))" % (a, vec(rng, depth - 1))
    if k == 8:
        return "Atan(%s)" % a
    if k == 9:
        return "Minus(%s)" % a
    return "(%s*%s)" % (a, rng.choice(["IntCst(2)", "IntCst(-3)", "RatCst(1,3)", P, B]))


def sca(rng, depth):
    """a random dimension-1 expression"""
    if depth <= 0 or rng.random() < 0.2:
        return rng.choice([B, P, "SqDist(%s,%s)" % (X, Y), "(%s|%s)" % (X, Y), "IntCst(2)", "Elem(%s,1)" % X])
    k = rng.randrange(14)
    if k == 0:
        return "Sum(%s)" % vec(rng, depth - 1)
    if k == 1:
        return "SqNorm2(%s)" % vec(rng, depth - 1)
    if k == 2:
        return "(%s|%s)" % (vec(rng, depth - 1), vec(rng, depth - 1))
    if k == 3:
        return "Exp(Minus(Square(%s)))" % sca(rng, depth - 1)
    if k == 4:
        return "Log(IntCst(1)+Square(%s))" % sca(rng, depth - 1)
    if k == 5:
        return "Sqrt(IntCst(1)+Square(%s))" % sca(rng, depth - 1)
    if k == 6:
        return "Rsqrt(IntCst(2)+Square(%s))" % sca(rng, depth - 1)
    if k == 7:
        return "Inv(IntCst(1)+Square(%s))" % sca(rng, depth - 1)
    if k == 8:
        return "(%s*%s)" % (sca(rng, depth - 1), sca(rng, depth - 1))
    if k == 9:
        return "(%s+%s)" % (sca(rng, depth - 1), sca(rng, depth - 1))
    if k == 10:
        return "(%s-%s)" % (sca(rng, depth - 1), sca(rng, depth - 1))
    if k == 11:
        return "Pow(%s,3)" % sca(rng, depth - 1)
    if k == 12:
        return "Norm2(%s+IntCst(2))" % vec(rng, depth - 1)
    return "(%s/(IntCst(1)+Square(%s)))" % (sca(rng, depth - 1), sca(rng, depth - 1))


def data(seed):
    g = np.random.default_rng(seed)
    return [g.uniform(-1, 1, (7, 3)), g.uniform(-1, 1, (9, 3)), g.uniform(0.5, 1.5, (9, 1)), g.uniform(0.5, 1.5, (1,))]


def close(a, b, what):
    scale = max(1.0, float(np.abs(b).max()))
    assert a.shape == b.shape and np.abs(a - b).max() <= 1e-9 * scale, (what, float(np.abs(a - b).max()), scale)


@pytest.mark.parametrize("seed", range(60))
def test_random_formula_forward_and_gradients(seed):
    rng = random.Random(1000 + seed)
    dim, F = (3, vec(rng, 3)) if seed % 2 else (1, sca(rng, 3))
    args = data(seed)
    # half of the seeds go through the ``use_fast_math`` lowering (products re-associated so that hoistable factors meet,
    # reciprocal + multiply for divisions: what float32 kernels are generated from); evaluated in float64 both agree to rounding
    fast = seed % 4 < 2
    for tagI in (0, 1):
        fwd = "Sum_Reduction(%s,%d)" % (F, tagI)
        out = ko.genred(fwd, [], args, "float64", sizes=(7, 9))
        close(dag_eval.evaluate(fwd, [], args, "float64", fast=fast, sizes=(7, 9)), out, fwd)
        # gradients of the Sum reduction with respect to x_i, y_j and the parameter, upstream gradient eta
        eta = np.random.default_rng(seed + 7).uniform(-1, 1, out.shape)
        for var in (X, Y, P, B):
            g = "Grad_WithSavedForward(%s,%s,Var(4,%d,%d),Var(5,%d,%d))" % (fwd, var, dim, tagI, dim, tagI)
            try:
                ref = ko.genred(g, [], args + [eta, out], "float64", sizes=(7, 9))
            except ZeroDivisionError:
                continue
            close(dag_eval.evaluate(g, [], args + [eta, out], "float64", fast=fast, sizes=(7, 9)), ref, g)
    if dim == 1:
        for red in ("Max_SumShiftExp_Reduction(%s,0)", "Min_Reduction(%s,1)", "Max_ArgMax_Reduction(%s,0)", "KMin_Reduction(%s,3,0)"):
            f = red % F
            close(dag_eval.evaluate(f, [], args, "float64", sizes=(7, 9)), ko.genred(f, [], args, "float64", sizes=(7, 9)), f)
