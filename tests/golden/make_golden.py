"""Generate golden input/output vectors by running the REAL reference (getkeops/keops v2.3, CPU
backend ``CpuReduc``) in the build container.  The reference cannot travel to the GPU box, so the
vectors are committed under ``tests/golden/cases/`` together with this script.

Run (from the repo root, in the container where /root/reference is mounted)::

    CXX=/usr/bin/g++ KEOPS_CACHE_FOLDER=/tmp/keops_cache \
    PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops python tests/golden/make_golden.py

Each case stores: the reduction string exactly as ``pykeops`` hands it to its binder
(``keops_binder[...](..., formula, aliases, nargs, dtype, 'torch', optional_flags)``,
``pykeops/pykeops/torch/generic/generic_red.py:91-103``), the input arrays, and the *raw* binder output
(before ``postprocess``).  Inputs are drawn from a fixed-seed numpy generator.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "cases")

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
C4 = "Max_SumShiftExp_Reduction((Var(0,1,1) - ((Sum(Square((Var(1,3,0) - Var(2,3,1)))) / IntCst(2)) / Var(3,1,2))),0)"


def V(k, d, c):
    return "Var(%d,%d,%d)" % (k, d, c)


# name, formula, [(dim, cat, kind)], dtype, (nx, ny)
#   kind: 'u' uniform[0,1), 'n' normal, 'p' positive uniform[0.5,2), 'eps' small positive scalar
CASES = []


def case(name, formula, args, dtypes=("float32", "float64"), sizes=(57, 83), aliases=None):
    for dt in dtypes:
        CASES.append(dict(name="%s_%s" % (name, dt), formula=formula, args=args, dtype=dt, sizes=sizes,
                          aliases=aliases))


XYB = [(3, 0, "u"), (3, 1, "u"), (1, 1, "u")]
case("gauss_sum", G, XYB)
case("gauss_grad_x", "Grad_WithSavedForward(%s,%s,%s,%s)" % (G, V(0, 3, 0), V(3, 1, 0), V(4, 1, 0)),
     XYB + [(1, 0, "n"), (1, 0, "n")])
case("gauss_grad_y", "Grad_WithSavedForward(%s,%s,%s,%s)" % (G, V(1, 3, 1), V(3, 1, 0), V(4, 1, 0)),
     XYB + [(1, 0, "n"), (1, 0, "n")])
case("gauss_grad_b", "Grad_WithSavedForward(%s,%s,%s,%s)" % (G, V(2, 1, 1), V(3, 1, 0), V(4, 1, 0)),
     XYB + [(1, 0, "n"), (1, 0, "n")])
case("gauss_sum_axis0", G[:-3] + ",1)", XYB)
case("gauss_randn_200x300", G, [(3, 0, "n"), (3, 1, "n"), (1, 1, "n")], sizes=(200, 300))
case("gauss_vec_b", "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,5,1)),0)",
     [(3, 0, "u"), (3, 1, "u"), (5, 1, "n")])
for K, D in ((3, 3), (10, 3), (10, 64), (1, 2)):
    case("argkmin_K%d_D%d" % (K, D),
         "ArgKMin_Reduction(Sum(Square((Var(0,%d,0) - Var(1,%d,1)))),%d,0)" % (D, D, K),
         [(D, 0, "n"), (D, 1, "n")], sizes=(64, 301))
case("argkmin_axis0", "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),4,1)", [(3, 0, "n"), (3, 1, "n")],
     sizes=(151, 40))
case("kmin", "KMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),5,0)", [(3, 0, "n"), (3, 1, "n")])
case("kmin_argkmin", "KMin_ArgKMin_Reduction(Sqrt(Sum(Square((Var(0,3,0) - Var(1,3,1))))),5,0)",
     [(3, 0, "n"), (3, 1, "n")])
case("kmin_argkmin_vec", "KMin_ArgKMin_Reduction(Square((Var(0,2,0) - Var(1,2,1))),3,0)", [(2, 0, "n"), (2, 1, "n")])
case("min", "Min_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),0)", [(3, 0, "n"), (3, 1, "n")])
case("argmin", "ArgMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),0)", [(3, 0, "n"), (3, 1, "n")])
case("min_argmin", "Min_ArgMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),0)", [(3, 0, "n"), (3, 1, "n")])
case("min_argmin_vec_axis0", "Min_ArgMin_Reduction((Var(0,3,0) - Var(1,3,1)),1)", [(3, 0, "n"), (3, 1, "n")])
case("max", "Max_Reduction((Var(0,3,0) * Var(1,3,1)),0)", [(3, 0, "n"), (3, 1, "n")])
case("argmax", "ArgMax_Reduction((Var(0,3,0) | Var(1,3,1)),0)", [(3, 0, "n"), (3, 1, "n")])
case("max_argmax", "Max_ArgMax_Reduction((Var(0,3,0) | Var(1,3,1)),1)", [(3, 0, "n"), (3, 1, "n")])
LSE_ARGS = [(1, 1, "u"), (3, 0, "u"), (3, 1, "u"), (1, 2, "eps")]
case("lse_softmin", C4, LSE_ARGS)
case("lse_grad_x", "Grad_WithSavedForward(%s,%s,%s,%s)" % (C4, V(1, 3, 0), V(4, 2, 0), V(5, 2, 0)),
     LSE_ARGS + [(2, 0, "n"), ("lse_fwd", 0, C4)])
case("lse_grad_f", "Grad_WithSavedForward(%s,%s,%s,%s)" % (C4, V(0, 1, 1), V(4, 2, 0), V(5, 2, 0)),
     LSE_ARGS + [(2, 0, "n"), ("lse_fwd", 0, C4)])
case("lse_grad_eps", "Grad_WithSavedForward(%s,%s,%s,%s)" % (C4, V(3, 1, 2), V(4, 2, 0), V(5, 2, 0)),
     LSE_ARGS + [(2, 0, "n"), ("lse_fwd", 0, C4)])
case("lse_weighted", "Max_SumShiftExpWeight_Reduction((Var(0,3,0) | Var(1,3,1)),0,Var(2,2,1))",
     [(3, 0, "n"), (3, 1, "n"), (2, 1, "n")])
case("sumsoftmaxweight", "Max_SumShiftExpWeight_Reduction((Var(0,3,0) | Var(1,3,1)),0,Concat(IntCst(1),Var(2,2,1)))",
     [(3, 0, "n"), (3, 1, "n"), (2, 1, "n")])
case("lse_axis0", "Max_SumShiftExp_Reduction(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1))))),1)", [(3, 0, "n"), (3, 1, "n")])
case("kernelsolve_matvec", "Sum_Reduction(Exp(-g*SqDist(x,y))*b,0)",
     [(16, 0, "u"), (16, 1, "u"), (1, 1, "n"), (1, 2, "p")], aliases=["x=Vi(16)", "y=Vj(16)", "b=Vj(1)", "g=Pm(1)"])
case("matern", "Sum_Reduction(Exp(-Sqrt(Sum(Square(x-y))))*(IntCst(1)+Sqrt(Sum(Square(x-y))))*b,0)",
     [(3, 0, "u"), (3, 1, "u"), (2, 1, "n")], aliases=["x=Vi(3)", "y=Vj(3)", "b=Vj(2)"])
OPS = {
    "inv_log": "Sum_Reduction(Log(IntCst(1)+Sum(Square(x-y)))*Inv(IntCst(2)+(x|y))*b,0)",
    "rsqrt_abs": "Sum_Reduction(Rsqrt(IntCst(1)+Sum(Abs(x-y)))*b,0)",
    "sign_relu_step": "Sum_Reduction(Sum(Sign(x-y)*ReLU(x-y)+Step(y-x))*b,0)",
    "ifelse": "Sum_Reduction(IfElse(x-y,x*y,x+y)*b,0)",
    "pow_powf": "Sum_Reduction(Sum(Pow(x-y,3))*b+Powf(IntCst(1)+Sum(Square(x-y)),Elem(y,0)),0)",
    "sincos": "Sum_Reduction(Sin(x|y)*Cos(Sum(x-y))*b,0)",
    "normalize_norm2": "Sum_Reduction((Normalize(x)|y)*Norm2(x-y)*b,0)",
    "concat_extract": "Sum_Reduction(Concat(Extract(x,1,2)*Elem(y,0),Sum(y)*b),0)",
    "sumT_elemT": "Sum_Reduction(SumT(b,3)*x+ElemT(Elem(y,2),3,1),0)",
    "matvec": "Sum_Reduction(MatVecMult(TensorProd(x,y),x)*b,0)",
    "clamp": "Sum_Reduction(Clamp(x-y,IntCst(0),b+IntCst(1)),0)",
    "div_scalar": "Sum_Reduction((x-y)/(IntCst(1)+Sum(Square(y)))*b,0)",
    "weightedsqdist": "Sum_Reduction(Exp(-WeightedSqDist(b+IntCst(1),x,y)),0)",
    "xlogx_sqrt": "Sum_Reduction(Sum(XLogX(IntCst(1)+Square(x-y)))*Sqrt(b*b),0)",
    "atan_round": "Sum_Reduction(Atan(x|y)*Round(IntCst(3)*b,0)+Floor(Sum(x)*b),0)",
}
for nm, f in OPS.items():
    ali = ["x=Vi(3)", "y=Vj(3)", "b=Vj(1)"]
    case("op_" + nm, f, [(3, 0, "u"), (3, 1, "u"), (1, 1, "p")], dtypes=("float32",), aliases=ali)
GRAD_OPS = ["inv_log", "rsqrt_abs", "ifelse", "pow_powf", "sincos", "normalize_norm2", "concat_extract", "matvec",
            "div_scalar", "weightedsqdist", "xlogx_sqrt", "clamp"]
for nm in GRAD_OPS:
    f = OPS[nm]
    ali = ["x=Vi(3)", "y=Vj(3)", "b=Vj(1)"]
    for vname, vstr in (("x", V(0, 3, 0)), ("y", V(1, 3, 1)), ("b", V(2, 1, 1))):
        if (nm, vname) == ("ifelse", "b"):
            # reference quirk: IfElse.DiffT (formulas/maths/IfElse.py) broadcasts the gradient to the
            # dimension of the condition, so d/db comes back with dim 3 instead of 1 -- not a parity target
            continue
        case("grad_%s_%s" % (nm, vname), "Grad_WithSavedForward(%s,%s,%s,%s)" % (f, vstr, "Var(3,DIMOUT,0)", "Var(4,DIMOUT,0)"),
             [(3, 0, "u"), (3, 1, "u"), (1, 1, "p"), ("eta", 0, f), ("fwd", 0, f)], dtypes=("float64",),
             aliases=ali + ["Var(3,DIMOUT,0)", "Var(4,DIMOUT,0)"])


# ---- forward-mode differentials and the differential operators built on them (added later: these cases draw their inputs
# from a generator seeded by their own name, so that `make_golden.py fm_ dop_ op2_` reproduces them without touching the others)
G2 = "Sum_Reduction(Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,2,1),0)"
XYB2 = [(3, 0, "u"), (3, 1, "u"), (2, 1, "n")]
for vname, vstr, tspec in (("x", V(0, 3, 0), (3, 0, "n")), ("y", V(1, 3, 1), (3, 1, "n")), ("b", V(2, 2, 1), (2, 1, "n"))):
    case("fm_diff_gauss_" + vname, "Diff_WithSavedForward(%s,%s,Var(3,%d,%d),Var(4,2,0))" % (G2, vstr, tspec[0], tspec[1]),
         XYB2 + [tspec, ("fwd", 0, G2)], dtypes=("float64",))
case("fm_diff_lse_x", "Diff_WithSavedForward(%s,%s,Var(4,3,0),Var(5,2,0))" % (C4, V(1, 3, 0)),
     LSE_ARGS + [(3, 0, "n"), ("lse_fwd", 0, C4)], dtypes=("float64",))
DOPS = {
    "laplacian": "Sum_Reduction(Laplacian(Exp(-SqDist(x,y)),x)*b,0)",
    "divergence": "Sum_Reduction(Divergence(x*Exp(-SqDist(x,y)),x)*b,0)",
    "gradmatrix": "Sum_Reduction(GradMatrix((x-y)*SqDist(x,y),x)*b,0)",
    "factorize": "Sum_Reduction(Factorize(Exp(-SqDist(x,y))+Exp(-SqDist(x,y))*b,Exp(-SqDist(x,y))),0)",
    "autofactorize": "Sum_Reduction(AutoFactorize(Exp(-SqDist(x,y))*b+Exp(-SqDist(x,y))),0)",
    "formula_grad": "Sum_Reduction(Grad(Exp(-SqDist(x,y))*b,x,Var(3,1,0)),0)",
    "formula_diff": "Sum_Reduction(Diff(Exp(-SqDist(x,y))*b,x,Var(4,3,0)),0)",
}
for nm, f in DOPS.items():
    case("dop_" + nm, f, [(3, 0, "u"), (3, 1, "u"), (1, 1, "p"), (1, 0, "n"), (3, 0, "n")], dtypes=("float64",),
         aliases=["x=Vi(3)", "y=Vj(3)", "b=Vj(1)", "e=Vi(1)", "u=Vi(3)"])


OPS2 = {
    "asin_acos": "Sum_Reduction(Asin((x|y)/IntCst(4))*Acos(Elem(x,0)*Elem(y,1)/IntCst(2))*b,0)",
    "atan2": "Sum_Reduction(Atan2(x-y,x+y+IntCst(1))*b,0)",
    "argminmax_vec": "Sum_Reduction((ArgMax(x-y)+IntCst(3)*ArgMin(x*y)+Max(x+y)+Min(x-y))*b,0)",
    "clampint": "Sum_Reduction(ClampInt(IntCst(3)*(x-y),-1,1)*b,0)",
    "diffclampint": "Sum_Reduction(DiffClampInt(IntCst(3)*(x-y),-1,1)*b,0)",
    "compare": "Sum_Reduction((LessThan(x,y)+LessOrEqual(y,x)*IntCst(2)+Equal(Round(x,0),Round(y,0))*IntCst(4)"
               "+NotEqual(Floor(IntCst(2)*x),Floor(IntCst(2)*y))*IntCst(8))*b,0)",
    "extractT_cst": "Sum_Reduction(ExtractT(x,1,5)*b*IntInv(3)+RatCst(2,7),0)",
    "mod": "Sum_Reduction(Mod(IntCst(5)*(x-y),IntCst(2),IntCst(-1))*b,0)",
    "modvar": "Sum_Reduction(Mod(IntCst(5)*(x-y),y+IntCst(1),IntCst(-1))*b,0)",     # variable modulus: d/dn = -Floor((x-d)/n)
    "onehot": "Sum_Reduction(OneHot(Floor(IntCst(3)*b-IntCst(1)),5)*Elem(x,0),0)",
    "sinxdivx": "Sum_Reduction(SinXDivX(IntCst(4)*(x-y))*b,0)",
    "sqnorms": "Sum_Reduction(SqNorm2(x-y)*WeightedSqNorm(b+IntCst(1),x+y)*b,0)",
    "vecmatmult": "Sum_Reduction(VecMatMult(x,TensorProd(x,y))*b,0)",
    "zero": "Sum_Reduction(Zero(3)+x*b,0)",
}
for nm, f in OPS2.items():
    case("op2_" + nm, f, [(3, 0, "u"), (3, 1, "u"), (1, 1, "p")], aliases=["x=Vi(3)", "y=Vj(3)", "b=Vj(1)"])


GRAD_OPS2 = ["asin_acos", "atan2", "argminmax_vec", "clampint", "extractT_cst", "mod", "modvar", "sinxdivx", "sqnorms", "vecmatmult"]
for nm in GRAD_OPS2:
    f = OPS2[nm]
    ali = ["x=Vi(3)", "y=Vj(3)", "b=Vj(1)"]
    for vname, vstr in (("x", V(0, 3, 0)), ("y", V(1, 3, 1)), ("b", V(2, 1, 1))):
        case("op2grad_%s_%s" % (nm, vname), "Grad_WithSavedForward(%s,%s,%s,%s)" % (f, vstr, "Var(3,DIMOUT,0)", "Var(4,DIMOUT,0)"),
             [(3, 0, "u"), (3, 1, "u"), (1, 1, "p"), ("eta", 0, f), ("fwd", 0, f)], dtypes=("float64",),
             aliases=ali + ["Var(3,DIMOUT,0)", "Var(4,DIMOUT,0)"])


# ---- second-order gradients (what autograd builds for grad-of-grad, cf. test_lazytensor_gaussian_cpu.py) and the gradients
# of a weighted soft-max reduction
G1X = "Grad_WithSavedForward(%s,%s,Var(3,1,0),Var(4,1,0))" % (G, V(0, 3, 0))
G1Y = "Grad_WithSavedForward(%s,%s,Var(3,1,0),Var(4,1,0))" % (G, V(1, 3, 1))
A5 = XYB + [("eta", 0, G), ("fwd", 0, G)]
case("so_gauss_xx", "Grad_WithSavedForward(%s,%s,Var(5,3,0),Var(6,3,0))" % (G1X, V(0, 3, 0)),
     A5 + [("eta", 0, G1X, 5), ("fwd", 0, G1X, 5)], dtypes=("float64",))
case("so_gauss_xy", "Grad_WithSavedForward(%s,%s,Var(5,3,0),Var(6,3,0))" % (G1X, V(1, 3, 1)),
     A5 + [("eta", 0, G1X, 5), ("fwd", 0, G1X, 5)], dtypes=("float64",))
case("so_gauss_x_eta", "Grad_WithSavedForward(%s,%s,Var(5,3,0),Var(6,3,0))" % (G1X, V(3, 1, 0)),
     A5 + [("eta", 0, G1X, 5), ("fwd", 0, G1X, 5)], dtypes=("float64",))
case("so_gauss_yb", "Grad_WithSavedForward(%s,%s,Var(5,3,1),Var(6,3,1))" % (G1Y, V(2, 1, 1)),
     A5 + [("eta", 1, G1Y, 5), ("fwd", 1, G1Y, 5)], dtypes=("float64",))
SSW = "Max_SumShiftExpWeight_Reduction((Var(0,3,0) | Var(1,3,1)),0,Concat(IntCst(1),Var(2,2,1)))"
SSW_ARGS = [(3, 0, "n"), (3, 1, "n"), (2, 1, "n")]
for vname, vstr in (("x", V(0, 3, 0)), ("y", V(1, 3, 1)), ("b", V(2, 2, 1))):
    case("ssw_grad_" + vname, "Grad_WithSavedForward(%s,%s,Var(3,4,0),Var(4,4,0))" % (SSW, vstr),
         SSW_ARGS + [(4, 0, "n"), ("fwd", 0, SSW)], dtypes=("float64",))


# ---- exact ties: integer-valued points, so that many columns share the extremal value (the smallest index must win)
TI = [(3, 0, "i"), (3, 1, "i")]
case("ties_argmin", "ArgMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),0)", TI, sizes=(64, 301))
case("ties_min_argmin", "Min_ArgMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),0)", TI, sizes=(64, 301))
case("ties_max_argmax_axis0", "Max_ArgMax_Reduction((Var(0,3,0) | Var(1,3,1)),1)", TI, sizes=(151, 80))
case("ties_argkmin", "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),7,0)", TI, sizes=(64, 301))
case("ties_kmin_argkmin", "KMin_ArgKMin_Reduction(Sum(Abs((Var(0,3,0) - Var(1,3,1)))),5,1)", TI, sizes=(120, 90))


# ---- tensor contractions on the internal dimension (maths/TensorDot.py, maths/Kron.py): the matrix-vector special case of
# test_lazytensor_tensordot.py, the permuted 3-axis contraction of test_torch.py:680-730, the Kronecker products of test_kron.py
TD = {
    "matvec": ("TensorDot(x,y,[4,7],[7],[1],[0])", 28, 7),
    "matmat_w": ("TensorDot(x,y,[2,3],[3,2],[1],[0])*Exp(-SqDist(Extract(x,0,3),Extract(y,0,3)))", 6, 6),
    "permute": ("TensorDot(x,y,[2,3,2,2,4],[2,4,2,3,2,3],[4,0,2],[1,4,2],[4,3,2,0,1])", 96, 288),
    "outer_perm": ("TensorDot(x,y,[2,3],[2,2],[],[],[0,2,1,3])", 6, 4),
    "kron2": ("Kron(x,y,[2,3],[2,2])", 6, 4),
    "kron3": ("Kron(x,y,[2,3,2],[2,2,3])", 12, 12),
    "full": ("TensorDot(x,y,[3,4],[4,3],[0,1],[1,0])*x", 12, 12),
}
for nm, (f, dx, dy) in TD.items():
    ali = ["x=Vi(0,%d)" % dx, "y=Vj(1,%d)" % dy]
    red = "Sum_Reduction(%s,0)" % f
    case("td_" + nm, red, [(dx, 0, "n"), (dy, 1, "n")], aliases=ali, sizes=(23, 31))
    if nm in ("matmat_w", "permute", "kron2", "full"):
        for vname, vstr in (("x", V(0, dx, 0)), ("y", V(1, dy, 1))):
            case("td_grad_%s_%s" % (nm, vname), "Grad_WithSavedForward(%s,%s,%s,%s)" % (red, vstr, "Var(2,DIMOUT,0)", "Var(3,DIMOUT,0)"),
                 [(dx, 0, "n"), (dy, 1, "n"), ("eta", 0, red, 2), ("fwd", 0, red, 2)], dtypes=("float64",),
                 aliases=ali + ["Var(2,DIMOUT,0)", "Var(3,DIMOUT,0)"], sizes=(23, 31))


# ---- soft dynamic time warping (maths/SoftDTW_SqDist.py): value and gradients with respect to both series.
# The operands are expressions (2x, y+y), not bare variables: the reference's CPU backend reads bare variables straight from
# the argument arrays as ``(arg[0]+i*5)[k]``, and the loop counters ``int i`` / ``int j`` that SoftDTW_SqDist.py:46-75
# declares shadow the row / column indices in that expression, so on bare variables the CPU reference indexes out of its
# rows.  Expressions are evaluated into local arrays first, which is also what its CUDA backend does for every variable.
XS, YS = "(IntCst(2)*x)", "(y+y)"
SDTW = "Sum_Reduction(SoftDTW_SqDist(%s,%s,g)*b,0)" % (XS, YS)
SDTW_ALI = ["x=Vi(0,5)", "y=Vj(1,4)", "g=Pm(2,1)", "b=Vj(3,1)"]
SDTW_ARGS = [(5, 0, "n"), (4, 1, "n"), (1, 2, "p"), (1, 1, "p")]
case("sdtw_value", SDTW, SDTW_ARGS, aliases=SDTW_ALI, sizes=(23, 31))
case("sdtw_series_of_1", "Sum_Reduction(SoftDTW_SqDist(%s,Elem(%s,0),g)+SoftDTW_SqDist(Elem(%s,1),%s,g),0)" % (XS, YS, XS, YS),
     SDTW_ARGS[:3], aliases=SDTW_ALI[:3], sizes=(23, 31))
for vname, vstr in (("x", V(0, 5, 0)), ("y", V(1, 4, 1)), ("b", V(3, 1, 1))):
    case("sdtw_grad_" + vname, "Grad_WithSavedForward(%s,%s,Var(4,1,0),Var(5,1,0))" % (SDTW, vstr),
         SDTW_ARGS + [("eta", 0, SDTW, 4), ("fwd", 0, SDTW, 4)], dtypes=("float64",),
         aliases=SDTW_ALI + ["Var(4,1,0)", "Var(5,1,0)"], sizes=(23, 31))


# ---- B-spline basis functions (maths/BSpline.py): float32 only -- the reference's generator refuses double inputs (its
# ratios are C floats, "subtraction of two c_variable is only possible with same dtype"); one knot is repeated (empty span)
for order in (0, 1, 3):
    case("bspline_order%d" % order, "Sum_Reduction(BSpline(t,x-y,%d)*b,0)" % order,
         [(8, 2, "knots"), (1, 0, "u"), (1, 1, "u"), (1, 1, "p")], dtypes=("float32",),
         aliases=["t=Pm(0,8)", "x=Vi(1,1)", "y=Vj(2,1)", "b=Vj(3,1)"], sizes=(23, 31))


# ---- complex arithmetic on interleaved (re, im) pairs (formulas/complex/*.py): the non-uniform Fourier sum of
# test_complex.py, every operation once, and gradients with respect to complex and real (even-sized) variables
CPLX = {
    "nufft": "ComplexMult(z,ComplexExp1j(IntCst(-6)*p*f))",
    "exp_scal": "ComplexRealScal(f,ComplexExp(ComplexMult(z,w)))+ComplexRealScal(ComplexExp(Conj(w)),p)",
    "div_abs_angle": "Concat(ComplexDivide(z,w),Concat(ComplexAbs(z-w),ComplexAngle(ComplexAdd(z,ComplexSum(w)))))",
    "parts": "ComplexSubtract(Real2Complex(ComplexReal(z)*ComplexImag(w)),Imag2Complex(ComplexSquareAbs(ComplexMult(z,Elem(w,0)|Elem(w,1)*IntCst(0)+ComplexSumT(ComplexSum(w),4)))))",
    "sum_sumT": "ComplexMult(ComplexSum(z),w)+ComplexSumT(ComplexSum(ComplexMult(z,Conj(w))),4)",
}
CPLX["parts"] = "ComplexSubtract(Real2Complex(ComplexReal(z)*ComplexImag(w)),Imag2Complex(ComplexSquareAbs(ComplexMult(z,ComplexSumT(ComplexSum(w),4)))))"
CPLX_ALI = ["z=Vj(0,4)", "w=Vi(1,4)", "p=Vj(2,1)", "f=Vi(3,1)"]
CPLX_ARGS = [(4, 1, "n"), (4, 0, "n"), (1, 1, "u"), (1, 0, "u")]
for nm, f in CPLX.items():
    red = "Sum_Reduction(%s,0)" % f
    case("cplx_" + nm, red, CPLX_ARGS, aliases=CPLX_ALI, sizes=(23, 31))
    if nm in ("nufft", "exp_scal", "div_abs_angle", "sum_sumT"):
        for vname, vstr in (("z", V(0, 4, 1)), ("w", V(1, 4, 0))):
            if (nm, vname) == ("nufft", "w"):
                continue                    # w does not occur in that formula
            case("cplx_grad_%s_%s" % (nm, vname), "Grad_WithSavedForward(%s,%s,%s,%s)" % (red, vstr, "Var(4,DIMOUT,0)", "Var(5,DIMOUT,0)"),
                 CPLX_ARGS + [("eta", 0, red, 4), ("fwd", 0, red, 4)], dtypes=("float64",),
                 aliases=CPLX_ALI + ["Var(4,DIMOUT,0)", "Var(5,DIMOUT,0)"], sizes=(23, 31))


# ---- symbolic operators on formulas that are linear in a placeholder variable u = Var(4,3,0), which disappears from the
# result (formulas/LinearOperators/TraceOperator.py, AdjointOperator.py; LazyTensor.trace_operator)
LINOP_ALI = ["A=Vj(0,9)", "b=Vj(1,1)", "x=Vi(2,3)", "y=Vj(3,3)", "u=Vi(4,3)"]
LINOP_ARGS = [(9, 1, "n"), (1, 1, "p"), (3, 0, "n"), (3, 1, "n")]
LINOP = {
    "trace_diag": "TraceOperator((x*u)*b-u,u)*y+TraceOperator(u/(IntCst(1)+SqDist(x,y)),u)",
    "trace_weighted": "TraceOperator(u*y+SumT(Sum(u*x),3),u,x*y)",
    "trace_row": "TraceOperator((x|u)*b,u)+TraceOperator(Sum(u*y),u,y)",
    "trace_scalar_weight": "TraceOperator(u*Exp(-SqDist(x,y)),u,b)",
    "adjoint": "AdjointOperator((x*u)*b,u,y)+AdjointOperator((u|y)*x,u,x)+AdjointOperator(x*u-u/b,u,y)",
    "adjoint_row": "AdjointOperator(Sum(u*x)*b,u,b)+AdjointOperator(u*y+SumT(Sum(u*x),3),u,y)",
}
for nm, f in LINOP.items():
    case("linop_" + nm, "Sum_Reduction(%s,0)" % f, LINOP_ARGS, dtypes=("float64",), aliases=LINOP_ALI, sizes=(23, 31))


def flags(formula):
    head = formula.split("(")[0]
    sumlike = head in ("Sum_Reduction", "Max_SumShiftExp_Reduction", "Max_SumShiftExpWeight_Reduction",
                       "Grad_WithSavedForward")
    return dict(dtype_acc="auto", sum_scheme="block_sum" if sumlike else "direct_sum", use_fast_math=1,
                enable_chunks=1, multVar_highdim=0)


def run_reference(formula, aliases, args, dtype, nx, ny):
    import torch
    from pykeops.common.keops_io import keops_binder

    targs = [torch.from_numpy(np.ascontiguousarray(a)) for a in args]
    b = keops_binder["cpp"](0, 0, 0, False, -1, formula, list(aliases), len(targs), dtype, "torch", flags(formula))
    out = b.import_module().genred_pytorch(targs[0].device, None, nx, ny, 0, None, *targs)
    return out.numpy().copy(), int(b.dimout), int(b.tagIJ)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20260101)
    index = []
    only = sys.argv[1:]
    for c in CASES:
        if only and not any(o in c["name"] for o in only):
            continue
        if c["name"].startswith(("fm_", "dop_", "op2_", "op2grad_", "so_", "ssw_", "ties_", "td_", "sdtw_", "bspline_", "cplx_", "linop_")):
            import zlib
            rng = np.random.default_rng(zlib.crc32(c["name"].encode()))
        dt = np.dtype(c["dtype"])
        nx, ny = c["sizes"]
        args = []
        formula, aliases = c["formula"], c["aliases"]
        for spec in c["args"]:
            dim, cat, kind = spec[:3]
            nprev = spec[3] if len(spec) > 3 else None
            n = {0: nx, 1: ny}.get(cat)
            if dim in ("lse_fwd", "fwd", "eta"):
                # saved forward output (or a random upstream gradient of the same shape)
                take = nprev if nprev is not None else (4 if dim == "lse_fwd" else 3)
                fwd, dimout, _ = run_reference(kind, aliases[:take] if aliases else [], args[:take], c["dtype"], nx, ny)
                if dim == "eta":
                    a = rng.standard_normal(fwd.shape).astype(dt)
                else:
                    a = fwd.astype(dt)
                args.append(a)
                continue
            shape = (dim,) if cat == 2 else (n, dim)
            if kind == "u":
                a = rng.random(shape)
            elif kind == "n":
                a = rng.standard_normal(shape)
            elif kind == "p":
                a = 0.5 + 1.5 * rng.random(shape)
            elif kind == "eps":
                a = np.full(shape, 0.05)
            elif kind == "knots":
                a = np.sort(rng.random(shape) * 2.2 - 1.1)      # covers the range of x - y; one empty span
                a[..., 3] = a[..., 2]
            elif kind == "i":
                a = rng.integers(0, 3, shape).astype(np.float64)
            args.append(a.astype(dt))
        if "DIMOUT" in formula:
            dimout = args[-1].shape[1]
            formula = formula.replace("DIMOUT", str(dimout))
            aliases = [a.replace("DIMOUT", str(dimout)) for a in aliases]
        if aliases is None:
            aliases = []
        out, dimout, tagIJ = run_reference(formula, aliases, args, c["dtype"], nx, ny)
        path = os.path.join(OUT, c["name"] + ".npz")
        meta = dict(name=c["name"], formula=formula, aliases=aliases, dtype=c["dtype"], nx=nx, ny=ny,
                    nargs=len(args), dimout=dimout, tagIJ=tagIJ)
        np.savez_compressed(path, out=out, meta=json.dumps(meta), **{"arg%d" % k: a for k, a in enumerate(args)})
        index.append(meta)
        print("%-40s out %s  tagIJ=%d" % (c["name"], out.shape, tagIJ))
    # index.json lists every committed case (also after a partial regeneration)
    import glob
    metas = []
    for path in sorted(glob.glob(os.path.join(OUT, "*.npz"))):
        metas.append(json.loads(str(np.load(path)["meta"])))
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(metas, f, indent=1)
    print("wrote %d cases" % len(index))


if __name__ == "__main__":
    sys.exit(main())
