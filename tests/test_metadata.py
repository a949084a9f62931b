"""The metadata the engine derives from a reduction string (``LoadKeOpsB200.params``) against the tuple the real reference's
``keopscore.get_keops_dll`` returns for the same string (tagI, tagZero, dimred, dim, dimy, indsi, indsj, indsp, dimsx, dimsy,
dimsp -- ``keopscore/keopscore/get_keops_dll.py:78-159``), recorded for all golden formulas by tests/golden/make_metadata.py.
``get_keops_dll`` reports the variables in its *pre-swap* convention (for tagI == 1 the i / j lists are exchanged,
``utils/code_gen_utils.py:538-554``) and ``LoadKeOps.init`` swaps them back (``LoadKeOps.py:151-153``); the binder exposes the
post-swap lists, so the comparison undoes the swap."""
import json
import os

import pytest

from keops_b200.binder import get_binder
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "metadata.json")) as f:
    RECORDED = json.load(f)


@pytest.mark.parametrize("name", sorted(RECORDED))
def test_binder_metadata_matches_get_keops_dll(name):
    meta, _, _ = util.load_case(name)
    ref = RECORDED[name]
    p = get_binder(meta["formula"], meta["aliases"], meta["nargs"], meta["dtype"], util.default_flags(meta["formula"])).params
    assert p.tagI == ref["tagI"] and p.dim == ref["dim"] and p.dimred == ref["dimred"], (name, p.tagI, p.dim, p.dimred, ref)
    assert bool(p.tagZero) == bool(ref["tagZero"])
    swap = ref["tagI"] == 1
    indsi, indsj = (ref["indsj"], ref["indsi"]) if swap else (ref["indsi"], ref["indsj"])
    dimsx, dimsy = (ref["dimsy"], ref["dimsx"]) if swap else (ref["dimsx"], ref["dimsy"])
    if not ref["tagZero"]:
        assert list(p.indsi) == indsi and list(p.indsj) == indsj and list(p.indsp) == ref["indsp"], (name, p.indsi, p.indsj, ref)
        assert list(p.dimsx) == dimsx and list(p.dimsy) == dimsy and list(p.dimsp) == ref["dimsp"]
        assert p.dimy == ref["dimy"], (name, p.dimy, ref["dimy"])
