"""Random ``TensorDot`` configurations (shapes, contracted axis pairs, output permutations) on the REAL reference's CPU backend
vs this package's lowering (tests/dag_eval.py), float64: forward ``Sum_Reduction`` and formula-level gradients with respect to
both operands.  One g++ build of the reference per formula (~8 s): a script, not a test.

    CXX=/usr/bin/g++ KEOPS_CACHE_FOLDER=/tmp/keops_cache PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops \
        python experiments/random_tensordot_vs_reference.py [ntrials]
"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import dag_eval  # noqa: E402


def main(ntrials):
    from pykeops.numpy import Genred
    rng = random.Random(2026)
    bad = checked = 0
    for trial in range(ntrials):
        na, nb = rng.randint(1, 3), rng.randint(1, 3)
        da, db = [rng.randint(1, 3) for _ in range(na)], [rng.randint(1, 3) for _ in range(nb)]
        nc = rng.randint(0, min(na, nb))
        ca, cb = rng.sample(range(na), nc), rng.sample(range(nb), nc)
        for a_, b_ in zip(ca, cb):
            db[b_] = da[a_]
        perm = list(range(na + nb - 2 * nc))
        rng.shuffle(perm)
        Da, Db = int(np.prod(da)), int(np.prod(db))
        dout = Da * Db // (int(np.prod([da[c] for c in ca])) ** 2 if nc else 1)
        F = "TensorDot(Var(0,%d,0),Var(1,%d,1),%s,%s,%s,%s,%s)" % (Da, Db, da, db, ca, cb, perm)
        g = np.random.default_rng(trial)
        x, y, eta = g.standard_normal((5, Da)), g.standard_normal((6, Db)), g.standard_normal((5, dout))
        al = ["Var(0,%d,0)" % Da, "Var(1,%d,1)" % Db, "Var(2,%d,0)" % dout]
        cases = [("forward", F, [x, y], 0)]
        cases += [("grad_a", "Grad(%s,Var(0,%d,0),Var(2,%d,0))" % (F, Da, dout), [x, y, eta], 0),
                  ("grad_b", "Grad(%s,Var(1,%d,1),Var(2,%d,0))" % (F, Db, dout), [x, y, eta], 1)]
        for what, formula, a, axis in cases:
            try:
                ref = Genred(formula, al[:len(a)], reduction_op="Sum", axis=1 - axis)(*a, backend="CPU")
            except Exception as e:
                print("reference error", trial, what, type(e).__name__, str(e)[:100])
                continue
            got = dag_eval.evaluate("Sum_Reduction(%s,%d)" % (formula, axis), [], a, "float64", sizes=(5, 6))
            checked += 1
            err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
            if got.shape != ref.shape or err > 1e-12:
                bad += 1
                print("MISMATCH", trial, what, err, formula)
    print("%d trials, %d comparisons, %d mismatches" % (ntrials, checked, bad))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
