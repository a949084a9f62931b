"""``Genred`` / ``KernelSolve`` constructors must hand the engine exactly what the reference's do: the reduction string
(``<Op>_Reduction(<formula>[,K],<tagI>[,<formula2>])``), the completed alias list and the axis -- recorded from the real
``pykeops.torch`` with tests/golden/make_genred_strings.py (``pykeops/pykeops/torch/generic/generic_red.py:386-520``,
``pykeops/pykeops/common/operations.py:7-43``, ``parse_type.py:134-172``).  No GPU needed."""
import importlib.util
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "genred_strings.json")) as f:
    RECORDED = json.load(f)
_spec = importlib.util.spec_from_file_location("make_genred_strings", os.path.join(HERE, "golden", "make_genred_strings.py"))
GEN = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(GEN)


@pytest.fixture(scope="module")
def ours():
    from keops_b200.torch import Genred, KernelSolve
    return GEN.build(Genred, KernelSolve)


@pytest.mark.parametrize("key", sorted(RECORDED))
def test_constructor_builds_the_reference_reduction_string(key, ours):
    ref, got = RECORDED[key], ours[key]
    strip = lambda s: s.replace(" ", "")
    assert strip(got["formula"]) == strip(ref["formula"]), (got["formula"], ref["formula"])
    assert [strip(a) for a in got["aliases"]] == [strip(a) for a in ref["aliases"]]
    assert got["axis"] == ref["axis"]
    for k in ("reduction_op", "varinvpos"):
        if k in ref:
            assert got[k] == ref[k]
