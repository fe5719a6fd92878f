"""The multi-process runner of the oracle (CPU baseline of bench.py's reference arm) reproduces the
serial oracle bit for bit."""
import numpy as np
import pytest

import dpgo_b200 as D
import parity
from oracle import dist_pgo as odist
from oracle import dpgo as odpgo


@pytest.mark.parametrize("loss,workers", [("trivial", 2), ("huber", 3)])
def test_parallel_star_equals_serial(loss, workers):
    g, _, X0 = D.grid3d(6, 6, 5, seed=7)
    meas = parity.to_measurements(g)
    opts = odpgo.Options(loss=loss, preconditioner="BlockJacobi")
    a = odist.run(meas, g.num_poses, 4, opts, X0, 6, "star")
    b = odist.run(meas, g.num_poses, 4, opts, X0, 6, "star", workers=workers)
    assert np.array_equal(a["X"], b["X"])
    assert a["trace"] == b["trace"]
    assert a["fobj_nodes"] == b["fobj_nodes"]
    assert a["refined"] == b["refined"]
