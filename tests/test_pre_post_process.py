"""``preprocess`` / ``postprocess`` (reduction-op renaming and output post-processing) against recordings of the reference's
own functions (``pykeops/pykeops/common/operations.py:7-87``; tests/golden/make_postprocess.py).  CPU tensors, no GPU."""
import importlib.util
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "pre_post_process.json")) as f:
    RECORDED = json.load(f)
_spec = importlib.util.spec_from_file_location("make_postprocess", os.path.join(HERE, "golden", "make_postprocess.py"))
GEN = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(GEN)


def test_pre_and_post_processing_match_the_reference():
    from keops_b200.common import postprocess, preprocess
    ours = GEN.build(preprocess, postprocess)
    assert len(ours["pre"]) == len(RECORDED["pre"]) and len(ours["post"]) == len(RECORDED["post"])
    for a, b in zip(ours["pre"], RECORDED["pre"]):
        assert a == b, (a, b)
    for a, b in zip(ours["post"], RECORDED["post"]):
        assert a["op"] == b["op"] and len(a["result"]) == len(b["result"]), (a["op"], b["op"])
        for ta, tb in zip(a["result"], b["result"]):
            assert ta["dtype"] == tb["dtype"] and ta["shape"] == tb["shape"], (a["op"], ta["dtype"], tb["dtype"], ta["shape"], tb["shape"])
            assert ta["data"] == pytest.approx(tb["data"], rel=1e-6, abs=1e-7), a["op"]
