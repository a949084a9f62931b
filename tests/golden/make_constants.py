"""Record what the REAL reference computes in float64 for formulas built around integer / rational constants, for
tests/test_frontend.py::test_constant_folding_matches_reference.  The reference folds rational constants in Add / Subtract /
Mult / Minus / Divide-by-constant only (keopscore/keopscore/formulas/maths/{Add,Subtract,Mult,Minus,Divide}.py) and emits a
folded RatCst float-rounded in every dtype (variables/RatCst.py:22-24), so WHERE folding happens is visible in the 8th
digit of fp64 results.  Same environment as make_golden.py.

    CXX=/usr/bin/g++ PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops python tests/golden/make_constants.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
X = "Var(0,1,0)"
FORMULAS = [
    "(IntCst(2)/(IntCst(1)+Square(IntCst(2))))*{x}", "Inv(IntCst(1)+Square(IntCst(2)))*{x}", "Sqrt(IntCst(1)+Square(IntCst(2)))*{x}",
    "Log(IntCst(1)+Square(Inv(IntCst(5))))*{x}", "(IntCst(2)*IntInv(5))*{x}", "{x}/IntCst(5)", "{x}*RatCst(2,5)", "{x}/RatCst(2,3)",
    "(IntCst(2)*RatCst(1,3))*{x}", "(RatCst(1,3)+IntCst(1))*{x}", "({x}*IntCst(2))/IntCst(3)", "IntCst(2)*(IntCst(3)*{x})",
    "Minus(RatCst(1,3))*{x}", "({x}+IntCst(1))-IntCst(1)/IntCst(3)", "Pow(IntCst(3),2)*{x}", "Exp(IntCst(1))*{x}",
    "(IntCst(1)/IntCst(3))*{x}", "{x}*(IntCst(1)/IntCst(3))", "IntCst(2)/IntCst(3)+{x}", "Sqrt(IntCst(2))*{x}", "IntCst(1)/{x}",
    "IntCst(3)/{x}", "RatCst(1,3)/{x}", "{x}/(IntCst(1)/IntCst(3))", "(RatCst(1,3)*RatCst(3,7))*{x}", "(RatCst(1,3)-RatCst(1,7))*{x}",
    "{x}*IntCst(3)*RatCst(1,7)", "Square(RatCst(1,3))*{x}", "Inv(RatCst(3,7))*{x}", "Inv(IntCst(3))*{x}", "Square(IntCst(3))/IntCst(7)*{x}",
    "({x}-RatCst(1,3))+RatCst(1,3)", "IntCst(2)*{x}+IntCst(3)*{x}", "{x}+{x}", "{x}*RatCst(1,3)+{x}", "Minus(IntCst(2)*{x})/IntCst(3)",
    "Exp(RatCst(1,3)*{x})", "Powf({x},RatCst(1,3))", "Rsqrt(IntCst(3))*{x}", "Abs(Minus(IntCst(3)))/IntCst(7)*{x}",
]


def main():
    from pykeops.numpy import Genred
    x = np.array([[1.0], [0.7], [1.9]])
    y = np.zeros((2, 1))
    rec = []
    for f in FORMULAS:
        f = f.format(x=X)
        try:
            out = Genred(f + "+Var(1,1,1)", ["Var(0,1,0)", "Var(1,1,1)"], reduction_op="Sum", axis=1)(x, y, backend="CPU")
        except AttributeError as e:        # the reference's own folding rules crash on RatCst (+|-) IntCst (Add.py:74, Subtract.py:72)
            rec.append(dict(formula=f, error=str(e)))
            continue
        rec.append(dict(formula=f, out=[float.hex(float(v)) for v in out.ravel()]))
    with open(os.path.join(HERE, "constants.json"), "w") as fh:
        json.dump(dict(x=x.ravel().tolist(), cases=rec), fh, indent=0)
    print("recorded", len(rec))


if __name__ == "__main__":
    main()
