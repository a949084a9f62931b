"""numpy interpreter for the code generator's scalar DAG -- TEST INFRASTRUCTURE (CPU).

Lets the CPU test-suite check parser + autodiff + lowering + sign/FMA rewriting (everything except the
CUDA text itself) against the oracle and the golden vectors without a GPU.  Never used by the product."""
import numpy as np

from keops_b200.codegen.emit import lower_reduction
from keops_b200.formulas.parser import parse_reduction


def _fn(op, a, dt):
    one, zero = dt.type(1), dt.type(0)
    f = {
        "add": lambda: a[0] + a[1], "sub": lambda: a[0] - a[1], "mul": lambda: a[0] * a[1],
        "fma": lambda: a[0] * a[1] + a[2], "neg": lambda: -a[0], "rcp": lambda: one / a[0], "div": lambda: a[0] / a[1],
        "exp": lambda: np.exp(a[0]), "ex2": lambda: np.exp2(a[0]), "ex2n": lambda: np.exp2(-a[0]), "keep": lambda: a[0],
        "sqrt_rn": lambda: np.sqrt(a[0]), "log": lambda: np.log(a[0]), "lg2": lambda: np.log2(a[0]),
        "sqrt": lambda: np.sqrt(a[0]), "rsqrt": lambda: np.where(a[0] == 0, zero, one / np.sqrt(np.where(a[0] == 0, one, a[0]))),
        "sin": lambda: np.sin(a[0]), "cos": lambda: np.cos(a[0]), "asin": lambda: np.arcsin(a[0]), "acos": lambda: np.arccos(a[0]),
        "atan": lambda: np.arctan(a[0]), "atan2": lambda: np.arctan2(a[0], a[1]), "abs": lambda: np.abs(a[0]),
        "sign": lambda: np.sign(a[0]), "step": lambda: np.where(a[0] < 0, zero, one), "relu": lambda: np.where(a[0] < 0, zero, a[0]),
        "floor": lambda: np.floor(a[0]), "round0": lambda: np.where(a[0] >= 0, np.floor(a[0] + 0.5), np.ceil(a[0] - 0.5)),
        "xlogx": lambda: np.where(a[0] != 0, a[0] * np.log(np.where(a[0] != 0, a[0], one)), zero),
        "sinxdivx": lambda: np.where(a[0] != 0, np.sin(a[0]) / np.where(a[0] != 0, a[0], one), one),
        "powf": lambda: np.power(a[0].astype(np.float32), a[1].astype(np.float32)).astype(dt),
        "fmax": lambda: np.maximum(a[0], a[1]), "fmin": lambda: np.minimum(a[0], a[1]),
        "lt": lambda: (a[0] < a[1]).astype(dt), "le": lambda: (a[0] <= a[1]).astype(dt),
        "eq": lambda: (a[0] == a[1]).astype(dt), "ne": lambda: (a[0] != a[1]).astype(dt),
        "sel_ge0": lambda: np.where(a[0] >= 0, a[1], a[2]), "clamp": lambda: np.where(a[0] < a[1], a[1], np.where(a[0] > a[2], a[2], a[0])),
        "inrange": lambda: np.where(a[0] < a[1], zero, np.where(a[0] > a[2], zero, one)),
        "mod": lambda: a[0] - a[1] * np.floor((a[0] - a[2]) / a[1]),
    }
    return f[op]()


def evaluate(formula, aliases, args, dtype="float64", fast=False, sizes=None, prescale=False):
    """Dense evaluation of the *lowered* reduction: returns the raw output (nrow, dim).  ``prescale=True`` evaluates the
    pre-scaled variant (``emit._prescale_rewrite``) the way ``kb200_pack`` + ``kb200_reduce`` do: common shift = mean of
    the streamed distance coordinates, scaled copies ``(y - shift) * sigma`` rounded to ``dtype``."""
    red = parse_reduction(formula, aliases)
    dt = np.dtype(dtype)
    tagI = red.tagI
    low = lower_reduction(red, fast, prescale=prescale)
    if prescale and low is None:
        return None
    vs = {v.params[0]: (v.params[1], v.params[2]) for v in red.variables()}
    nrow = ncol = 1
    if sizes is not None:                      # (nx, ny): a side without variables still has a size
        nrow, ncol = (sizes[0], sizes[1]) if tagI == 0 else (sizes[1], sizes[0])
    for ind, (dim, cat) in vs.items():
        if cat == tagI:
            nrow = np.asarray(args[ind]).reshape(-1, dim).shape[0]
        elif cat == 1 - tagI:
            ncol = np.asarray(args[ind]).reshape(-1, dim).shape[0]
    if low is None or red.kind == "Zero":
        return np.zeros((nrow, red.dim), dt)
    val = {}
    shift = {}
    if prescale:
        for k, (ix, cx, iy, cy) in enumerate(low.prescale["pairs"]):
            shift[k] = np.asarray(args[iy], dtype=dt).reshape(ncol, -1)[:, cy].mean(dtype=np.float64).astype(dt)
    def ev_p(n):
        """A constant / parameter-only node, whatever its place in the evaluation order."""
        if n.uid not in val:
            if n.op == "const":
                val[n.uid] = np.full((1, 1), n.val, dt)
            elif n.op == "load":
                val[n.uid] = np.asarray(args[n.val[0]], dtype=dt).reshape(-1)[n.val[1]].reshape(1, 1)
            else:
                val[n.uid] = np.asarray(_fn(n.op, [ev_p(x) for x in n.args], dt), dtype=dt)
        return val[n.uid]

    for n in low.order:
        if n.op == "const":
            val[n.uid] = np.full((1, 1), n.val, dt)
        elif n.op == "load":
            ind, comp, cat = n.val
            if cat == 6:
                val[n.uid] = np.full((1, 1), shift[ind], dt)
                continue
            if cat == 7:
                k = [(q[2], q[3]) for q in low.prescale["pairs"]].index((ind, comp))
                sig = ev_p(low.prescale["sigma"])
                col = np.asarray(args[ind], dtype=dt).reshape(ncol, -1)[:, comp].reshape(1, ncol)
                val[n.uid] = ((col - shift[k]).astype(dt) * sig).astype(dt)
                continue
            a = np.asarray(args[ind], dtype=dt)
            if cat == 2:
                val[n.uid] = a.reshape(-1)[comp].reshape(1, 1)
            elif cat == tagI:
                val[n.uid] = a.reshape(nrow, -1)[:, comp].reshape(nrow, 1)
            else:
                val[n.uid] = a.reshape(ncol, -1)[:, comp].reshape(1, ncol)
        else:
            with np.errstate(all="ignore"):
                val[n.uid] = np.asarray(_fn(n.op, [val[x.uid] for x in n.args], dt), dtype=dt)
    full = lambda v: np.broadcast_to(v, (nrow, ncol))
    d = red.dimF
    if red.kind in ("Sum", "LSE"):
        terms = []
        scales = low.row_scales or [None] * len(low.acc_terms)
        for (mode, nodes, neg), h in zip(low.acc_terms, scales):
            if mode == "zero":
                terms.append(np.zeros((nrow, ncol), dt))
                continue
            t = full(val[nodes[0].uid] * val[nodes[1].uid]) if mode == "fma" else full(val[nodes[0].uid])
            if h is not None:       # row-side factor applied at write-back
                t = t * full(val[h.uid])
            terms.append(-t if neg else t)
        if red.kind == "Sum":
            return np.stack([t.sum(axis=1, dtype=np.float64) for t in terms], axis=1).astype(dt)
        F = full(val[low.value_nodes[0].uid])
        m = F.max(axis=1, keepdims=True)
        e = np.exp(F - m)
        return np.concatenate([m] + [(e * t).sum(axis=1, keepdims=True, dtype=np.float64) for t in terms], axis=1).astype(dt)
    Fv = np.stack([full(val[n.uid]) for n in low.value_nodes], axis=2)   # (nrow, ncol, d)
    K = red.K
    if red.kind in ("Min", "Max", "ArgMin", "ArgMax", "Min_ArgMin", "Max_ArgMax"):
        idx = Fv.argmin(axis=1) if "Min" in red.kind else Fv.argmax(axis=1)
        v = np.take_along_axis(Fv, idx[:, None, :], axis=1)[:, 0, :]
        if red.kind in ("Min", "Max"):
            return v
        if red.kind in ("ArgMin", "ArgMax"):
            return idx.astype(dt)
        return np.concatenate([v, idx.astype(dt)], axis=1)
    order = np.argsort(Fv, axis=1, kind="stable")[:, :K, :]
    vals = np.take_along_axis(Fv, order, axis=1)
    if red.kind == "KMin":
        return vals.reshape(nrow, K * d)
    if red.kind == "ArgKMin":
        return order.astype(dt).reshape(nrow, K * d)
    return np.stack([vals, order.astype(dt)], axis=2).reshape(nrow, 2 * K * d)
