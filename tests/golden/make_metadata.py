"""Record the metadata tuple the REAL reference computes for every golden formula (``keopscore.get_keops_dll``:
tagI, tagZero, dimred, dim, dimy, indsi, indsj, indsp, dimsx, dimsy, dimsp -- keopscore/keopscore/get_keops_dll.py:78-159,
documented :1-45) with the CPU map-reduce scheme, for tests/test_metadata.py.  Uses the compilation cache left by
make_golden.py (same formulas, same flags), so nothing is recompiled.  Same environment as make_golden.py."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def normalise(aliases):
    """what LoadKeOps.init does before calling get_keops_dll (pykeops/pykeops/common/keops_io/LoadKeOps.py:47-64)"""
    out = []
    for k, alias in enumerate(aliases):
        alias = alias.replace(" ", "")
        if "=" in alias:
            varname, var = alias.split("=")
            if var.startswith("Var("):
                out.append(alias)
                continue
            cat = 0 if "Vi" in var else (1 if "Vj" in var else 2)
            a = var[3:-1].split(",")
            ind, dim = (k, int(a[0])) if len(a) == 1 else (int(a[0]), int(a[1]))
            out.append("%s=Var(%d,%d,%d)" % (varname, ind, dim, cat))
    return out


def main():
    from tests import util
    from keopscore.get_keops_dll import get_keops_dll
    from pykeops.common.parse_type import get_optional_flags  # noqa: F401  (import check)
    rec = {}
    for name in util.golden_names():
        meta, args, _ = util.load_case(name)
        fl = util.default_flags(meta["formula"])
        dt = {"float32": "float", "float64": "double"}[meta["dtype"]]
        res = get_keops_dll("CpuReduc", meta["formula"], fl["enable_chunks"], -1, fl["multVar_highdim"], normalise(meta["aliases"]),
                            meta["nargs"], dt, dt, fl["sum_scheme"], 0, 0, 0, 0, fl["use_fast_math"], -1)
        (tag, srcfile, dllfile, tagI, tagZero, use_half, use_fast_math, cuda_block_size, use_chunk_mode, tag1D2D, dimred, dim,
         dimy, indsi, indsj, indsp, dimsx, dimsy, dimsp) = res[:19]
        rec[name] = dict(tagI=int(tagI), tagZero=int(tagZero), dimred=int(dimred), dim=int(dim), dimy=int(dimy),
                         indsi=[int(v) for v in indsi], indsj=[int(v) for v in indsj], indsp=[int(v) for v in indsp],
                         dimsx=[int(v) for v in dimsx], dimsy=[int(v) for v in dimsy], dimsp=[int(v) for v in dimsp])
    with open(os.path.join(HERE, "metadata.json"), "w") as f:
        json.dump(rec, f, indent=0)
    print("recorded %d entries" % len(rec))


if __name__ == "__main__":
    main()
