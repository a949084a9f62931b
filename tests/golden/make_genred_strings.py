"""Record what the REAL reference's ``pykeops.torch.Genred`` / ``KernelSolve`` constructors hand to the binder (reduction
string, completed aliases, axis) for a list of constructor calls (tests/test_genred_strings.py compares).  Run where
/root/reference is mounted, same environment as make_golden.py."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
AL = ["x = Vi(3)", "y = Vj(3)", "b = Vj(2)", "p = Pm(1)"]
CALLS = {
    "sum_axis0": dict(formula="Exp(-p*SqDist(x,y))*b", aliases=AL, reduction_op="Sum", axis=0),
    "sum_axis1": dict(formula="Exp(-p*SqDist(x,y))*b", aliases=AL, reduction_op="Sum", axis=1),
    "lse": dict(formula="-p*SqDist(x,y)", aliases=AL, reduction_op="LogSumExp", axis=1),
    "lse_weight": dict(formula="-p*SqDist(x,y)", aliases=AL, reduction_op="LogSumExp", axis=1, formula2="b"),
    "sumsoftmax": dict(formula="-p*SqDist(x,y)", aliases=AL, reduction_op="SumSoftMaxWeight", axis=1, formula2="b"),
    "min": dict(formula="SqDist(x,y)", aliases=AL, reduction_op="Min", axis=1),
    "argmin": dict(formula="SqDist(x,y)", aliases=AL, reduction_op="ArgMin", axis=0),
    "min_argmin": dict(formula="SqDist(x,y)", aliases=AL, reduction_op="Min_ArgMin", axis=1),
    "max_argmax": dict(formula="(x|y)", aliases=AL, reduction_op="Max_ArgMax", axis=1),
    "kmin": dict(formula="SqDist(x,y)", aliases=AL, reduction_op="KMin", axis=1, opt_arg=3),
    "argkmin": dict(formula="SqDist(x,y)", aliases=AL, reduction_op="ArgKMin", axis=1, opt_arg=5),
    "kmin_argkmin": dict(formula="SqDist(x,y)", aliases=AL, reduction_op="KMin_ArgKMin", axis=0, opt_arg=2),
    "positional_aliases": dict(formula="Exp(-SqDist(Var(0,3,0),Var(1,3,1)))*Var(2,1,1)", aliases=["Var(0,3,0)", "Var(1,3,1)", "Var(2,1,1)"],
                               reduction_op="Sum", axis=1),
    "explicit_positions": dict(formula="Square(p-a)*Exp(x+y)", aliases=["p=Pm(0,1)", "a=Vj(1,1)", "x=Vi(2,3)", "y=Vj(3,3)"],
                               reduction_op="Sum", axis=1),
    "incomplete_aliases": dict(formula="Exp(-SqDist(x,Var(1,3,1)))*Var(2,1,1)", aliases=["x=Vi(0,3)"], reduction_op="Sum", axis=1),
}
SOLVES = {
    "solve_axis0": dict(formula="Exp(-p*SqDist(x,y))*b", aliases=AL, varinvalias="b", axis=0),
    "solve_axis1": dict(formula="Exp(-p*SqDist(x,y))*b", aliases=AL, varinvalias="b", axis=1),
}


def build(Genred, KernelSolve):
    out = {}
    for name, kw in CALLS.items():
        g = Genred(**kw)
        out["genred:" + name] = dict(formula=g.formula, aliases=list(g.aliases), axis=g.axis, reduction_op=g.reduction_op)
    for name, kw in SOLVES.items():
        k = KernelSolve(**kw)
        out["solve:" + name] = dict(formula=k.formula, aliases=list(k.aliases), axis=k.axis, varinvpos=k.varinvpos)
    return out


if __name__ == "__main__":
    from pykeops.torch import Genred, KernelSolve
    rec = build(Genred, KernelSolve)
    with open(os.path.join(HERE, "genred_strings.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("recorded %d entries" % len(rec))
