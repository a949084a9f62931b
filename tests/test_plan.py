"""Launch planner invariants (keops_b200/runtime/plan.py): the column splits cover the streamed range exactly once, chunks
are whole tiles, workspaces stay under the limit, row segments tile the output rows."""
import itertools

import pytest

from keops_b200.runtime import plan


@pytest.mark.parametrize("nrow,ncol,TJ", list(itertools.product(
    [1, 1000, 10000, 125000, 1000000, 3000001], [1, 255, 4096, 1000000, 5000003], [256, 512, 1024])))
def test_launch_shape_covers_the_columns(nrow, ncol, TJ):
    rpb, slots = 1024, 148 * 4
    ns, chunk = plan.launch_shape(nrow, ncol, rpb, TJ, slots, 1, 4)
    assert ns >= 1 and chunk >= 1
    assert chunk % TJ == 0 or ns == 1
    assert (ns - 1) * chunk < ncol <= ns * chunk          # every split is non-empty, the last one reaches the end
    if ns > 1:
        assert ns * nrow * 1 * 4 <= (2 << 30)


@pytest.mark.parametrize("nrow", [10000, 125000, 250000, 999424, 1000000, 2500000])
def test_launch_segments_tile_the_rows(nrow):
    TJ, rpb, slots = 1024, 1024, 296
    ntiles = -(-1000000 // TJ)
    segs = plan.launch_segments(nrow, 1000000, rpb, TJ, slots, 20, 4, split_penalty_tiles=0.002 * ntiles)
    assert segs[0][0] == 0 and segs[-1][1] == nrow
    for (a0, a1, ns, chunk), nxt in zip(segs, segs[1:] + [None]):
        assert a0 < a1 and ns >= 1 and (ns - 1) * chunk < 1000000 <= ns * chunk
        if nxt is not None:
            assert a1 == nxt[0] and a1 % rpb == 0          # segments meet on a row-block boundary


def test_workspace_limit_caps_the_splits():
    ns, chunk = plan.launch_shape(1000, 10 ** 7, 1024, 256, 592, 20, 4, ws_limit_bytes=1000 * 20 * 4 * 3)
    assert ns <= 3
