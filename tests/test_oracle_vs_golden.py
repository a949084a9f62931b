"""CPU: pin the oracle (oracle/keops_oracle.py) against outputs of the real reference.

Golden vectors were produced by tests/golden/make_golden.py, which runs getkeops/keops v2.3's own
CPU backend; plus the reference's single hard-coded known answer
(pykeops/pykeops/torch/test_install.py:5-25: SqNorm2(x-y) summed -> [63., 90.])."""
import numpy as np
import pytest

from oracle import keops_oracle as ko
from tests import util


def test_reference_install_known_answer():
    x = np.arange(1, 10, dtype=np.float32).reshape(-1, 3)
    y = np.arange(3, 9, dtype=np.float32).reshape(-1, 3)
    out = ko.genred("Sum_Reduction(SqNorm2(x-y),1)", ["x=Vi(3)", "y=Vj(3)"], [x, y], "float32")
    assert out.ravel().tolist() == [63.0, 90.0]


@pytest.mark.parametrize("name", util.golden_names())
def test_oracle_matches_reference(name):
    meta, args, ref = util.load_case(name)
    try:
        got = ko.genred(meta["formula"], meta["aliases"], args, meta["dtype"], sizes=(meta["nx"], meta["ny"]))
    except (NotImplementedError, NameError) as e:
        pytest.skip("oracle does not restate this op: %s" % e)
    assert got.shape == ref.shape
    cols = util.mixed_index_columns(meta["formula"], ref.shape[1])
    if cols is None:  # KMin_ArgKMin: [K][2][dimF]
        K2 = ref.shape[1]
        d = ko.parse(meta["formula"], meta["aliases"]).F.dim
        cols = np.tile(np.r_[np.zeros(d, bool), np.ones(d, bool)], K2 // (2 * d))
    if cols.any():
        # indices: bit-exact (the oracle computes in the same dtype with the same operation order)
        assert np.array_equal(got[:, cols], ref[:, cols]), name
    if (~cols).any():
        # the CPU reference itself is fp32 with libm exp: 1e-5 relative is the north-star bar
        # element-wise against the condition scale sum_j |F_ij| (tests/util.py::assert_close)
        A = ko.abs_scale(meta["formula"], meta["aliases"], args, "float64", sizes=(meta["nx"], meta["ny"]))
        util.assert_close(got[:, ~cols], ref[:, ~cols], meta["dtype"], name, rtol=util.rtol_for(meta), scale=A[:, ~cols],
                          floor=util.noise_floor(args, meta["dtype"], meta["ny"] if meta["tagIJ"] == 0 else meta["nx"]))
