"""Record the REAL reference's ``preprocess`` / ``postprocess`` (pykeops/pykeops/common/operations.py:7-87) on fixed raw
engine outputs, for tests/test_pre_post_process.py.  Run where /root/reference is mounted (same environment as
make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PRE = [("Sum", None), ("SumSoftMaxWeight", "b"), ("SoftMax", "b"), ("LogSumExp", None), ("LogSumExp", "b"),
       ("LogSumExpWeight", "b"), ("Min", None), ("ArgKMin", None), ("Max_ArgMax", None)]
# reduction_op, opt_arg, trailing dimension of the raw output
POST = [("Sum", None, 3), ("SumSoftMaxWeight", None, 5), ("ArgMin", None, 2), ("ArgMax", None, 1), ("Min_ArgMin", None, 6),
        ("Max_ArgMax", None, 2), ("KMin", 3, 6), ("KMin", 4, 4), ("ArgKMin", 3, 6), ("ArgKMin", 5, 5), ("KMin_ArgKMin", 2, 12),
        ("KMin_ArgKMin", 3, 6), ("LogSumExp", None, 2), ("LogSumExp", None, 4), ("Min", None, 3)]


def raw(op, dim, lead):
    rng = np.random.default_rng(abs(hash((op, dim, lead))) % (2 ** 32) if False else (len(op) * 1000 + dim * 10 + len(lead)))
    a = rng.random(lead + (dim,)) * 9.0 + 0.5
    if "Arg" in op or op in ("KMin_ArgKMin",):
        a = np.floor(a)                      # index-like values must survive the cast to integers
    return a.astype(np.float32)


def build(preprocess, postprocess):
    import torch
    out = {"pre": [], "post": []}
    for op, f2 in PRE:
        try:
            r = preprocess(op, f2)
            out["pre"].append(dict(op=op, formula2=f2, result=list(r)))
        except ValueError as e:
            out["pre"].append(dict(op=op, formula2=f2, error=True))
    try:
        preprocess("LogSumExpWeight", None)
        out["pre"].append(dict(op="LogSumExpWeight", formula2=None, result="no error"))
    except ValueError:
        out["pre"].append(dict(op="LogSumExpWeight", formula2=None, error=True))
    for op, opt, dim in POST:
        for lead in ((7,), (2, 7)):
            a = torch.from_numpy(raw(op, dim, lead))
            r = postprocess(a, op, opt)
            rs = r if isinstance(r, tuple) else (r,)
            out["post"].append(dict(op=op, opt_arg=opt, dim=dim, lead=list(lead),
                                    result=[dict(dtype=str(t.dtype), shape=list(t.shape), data=t.double().flatten().tolist()) for t in rs]))
    return out


if __name__ == "__main__":
    from pykeops.common.operations import preprocess, postprocess
    rec = build(preprocess, lambda a, op, opt: postprocess(a, "torch", op, None, opt, "float32"))
    with open(os.path.join(HERE, "pre_post_process.json"), "w") as f:
        json.dump(rec, f)
    print("recorded %d + %d entries" % (len(rec["pre"]), len(rec["post"])))
