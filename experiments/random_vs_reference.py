"""Randomized hunt against the REAL reference (CPU backend): the seeded random formulas of tests/test_random_formulas.py,
forward (`Sum_Reduction`) and formula-level gradients with respect to the i-variable, evaluated by the unmodified
pykeops.numpy on CPU and by this package's lowering (tests/dag_eval.py) in float64.  Each new formula costs the reference
one g++ build (~8 s), which is why this is a script and not part of the test-suite.

    CXX=/usr/bin/g++ KEOPS_CACHE_FOLDER=/tmp/keops_cache PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops \
        python experiments/random_vs_reference.py [first_seed last_seed]

Round 1: seeds 0-39 forward and 40-69 gradients -> no discrepancy above 1e-12 / 1e-11.
"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import dag_eval, test_random_formulas as T  # noqa: E402

PAD = "+IntCst(0)*(Var(0,3,0)|Var(1,3,1))*Var(2,1,1)*Var(3,1,2)"     # makes every variable appear, contributes exactly 0


def main(first, last):
    from pykeops.numpy import Genred
    bad = 0
    for seed in range(first, last):
        rng = random.Random(1000 + seed)
        dim, F = (3, T.vec(rng, 3)) if seed % 2 else (1, T.sca(rng, 3))
        args = T.data(seed)
        eta = np.random.default_rng(seed + 7).uniform(-1, 1, (7, dim))
        al = ["Var(0,3,0)", "Var(1,3,1)", "Var(2,1,1)", "Var(3,1,2)", "Var(4,%d,0)" % dim]
        G = "Grad(%s,Var(0,3,0),Var(4,%d,0))" % (F, dim)
        for what, formula, a, tol in (("forward", F, args, 1e-12), ("gradient", G, args + [eta], 1e-11)):
            try:
                ref = Genred(formula + PAD + ("*Sum(Var(4,%d,0))" % dim if what == "gradient" else ""),
                             al[:len(a)], reduction_op="Sum", axis=1)(*a, backend="CPU")
            except Exception as e:          # formulas the reference itself rejects
                print("reference error", seed, what, type(e).__name__, str(e)[:100])
                continue
            got = dag_eval.evaluate("Sum_Reduction(%s,0)" % formula, [], a, "float64", sizes=(7, 9))
            err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
            if err > tol:
                bad += 1
                print("MISMATCH", seed, what, err, formula)
    print("seeds %d-%d: %d mismatches" % (first, last - 1, bad))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 2 else 0, int(sys.argv[2]) + 1 if len(sys.argv) > 2 else 10)
