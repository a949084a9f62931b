"""keops_b200's LazyTensor must map every method / operator to the same KeOps operation tree as the reference front-end:
the reference's formula strings were recorded with tests/golden/make_lazytensor_strings.py (real pykeops.torch.LazyTensor,
symbolic variables) and are compared here after parsing both sides into the hash-consed IR -- identical trees are the very
same object.  No GPU needed: symbolic LazyTensors never touch data."""
import importlib.util
import json
import os

import pytest

from keops_b200.formulas.parser import parse_formula

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "lazytensor_formulas.json")) as f:
    RECORDED = json.load(f)
_spec = importlib.util.spec_from_file_location("make_lazytensor_strings", os.path.join(HERE, "golden", "make_lazytensor_strings.py"))
GEN = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(GEN)


@pytest.fixture(scope="module")
def ours():
    from keops_b200.torch import LazyTensor          # importing the torch front-end needs libkb200.so, not a GPU
    return GEN.build(LazyTensor)


def _norm(f):
    """python floats become parameter variables labelled by id(): relabel them 1000, 1001, ... in order of appearance"""
    import re
    seen = {}

    def sub(m):
        k = int(m.group(1))
        if k < 1000:
            return m.group(0)
        seen.setdefault(k, 1000 + len(seen))
        return "Var(%d," % seen[k]
    return re.sub(r"Var\((\d+),", sub, f)


def _same_tree(a, b):
    return parse_formula(_norm(a)) is parse_formula(_norm(b))


@pytest.mark.parametrize("key", sorted(k for k in RECORDED if k.startswith("expr:")))
def test_expression_maps_to_the_reference_operation_tree(key, ours):
    ref, got = RECORDED[key], ours[key]
    assert got["ndim"] == ref["ndim"], (got, ref)
    assert got["is_complex"] == ref["is_complex"] and got["shape"] == ref["shape"], (got, ref)
    assert _same_tree(got["formula"], ref["formula"]), "\nours: %s\nref:  %s" % (got["formula"], ref["formula"])


@pytest.mark.parametrize("key", sorted(k for k in RECORDED if k.startswith("red:")))
def test_reduction_hands_the_reference_arguments_to_the_engine(key, ours):
    ref, got = RECORDED[key], ours[key]
    assert got["reduction_op"] == ref["reduction_op"] and got["axis"] == ref["axis"], (got, ref)
    assert str(got["opt_arg"]) == str(ref["opt_arg"])
    assert _same_tree(got["formula"], ref["formula"])
    assert (got["formula2"] is None) == (ref["formula2"] is None)
    if ref["formula2"] is not None:
        assert _same_tree(got["formula2"], ref["formula2"])
