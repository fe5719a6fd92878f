"""The C++/OpenMP restated CPU baseline (oracle/cpu_dpgo.cpp -> oracle/_ref/libcpu_dpgo.so, timed by
bench.py --impl reference) against the numpy oracle: same per-node objective traces, same refinement
decisions, same final poses, for both drivers and all losses, with both threading modes."""
import numpy as np
import pytest

import dpgo_b200 as D
from oracle import cpu_ref
from oracle import dist_pgo as odist
from oracle import dpgo as odpgo
from parity import to_measurements

pytestmark = pytest.mark.skipif(not cpu_ref.available(), reason="oracle/_ref/libcpu_dpgo.so not built (needs /root/reference)")


def _both(g, nodes, X0, iters, alg, loss, mode=0, **kw):
    opts = odpgo.Options(loss=loss, preconditioner="BlockJacobi", **kw)
    meas = to_measurements(g)
    ref = odist.run(meas, g.num_poses, nodes, opts, X0, iters, alg, log_global=False)
    got = cpu_ref.run(meas, g.num_poses, nodes, opts, X0, iters, alg, workers=1, threads=4, mode=mode)
    return ref, got


@pytest.mark.parametrize("alg", ["star", "hash"])
@pytest.mark.parametrize("loss", ["trivial", "huber", "gm", "welsch"])
def test_cpu_baseline_matches_numpy_oracle_se3(alg, loss):
    g, _, X0 = D.grid3d(6, 6, 6, seed=1)
    ref, got = _both(g, 4, X0, 12, alg, loss)
    a, b = np.array(ref["fobj_nodes"]), np.array(got["fobj_nodes"])
    assert np.abs(a - b).max() <= 1e-9 * np.abs(a).max()
    assert np.array_equal(np.array(ref["refined"]), np.array(got["refined"]))
    assert np.abs(ref["X"] - got["X"]).max() < 1e-7


def test_cpu_baseline_se2_outliers_and_node_parallel_mode():
    g, _, X0 = D.city2d(14, 12, outlier_fraction=0.2, seed=9)
    ref, got = _both(g, 4, X0, 10, "hash", "gm", mode=1)
    a, b = np.array(ref["fobj_nodes"]), np.array(got["fobj_nodes"])
    assert np.abs(a - b).max() <= 1e-9 * np.abs(a).max()
    assert np.abs(ref["X"] - got["X"]).max() < 1e-7


def test_cpu_baseline_mm_scheme_and_jacobi():
    g, _, X0 = D.grid3d(6, 6, 6, seed=1)
    opts = odpgo.Options(loss="trivial", preconditioner="Jacobi", scheme="MM")
    meas = to_measurements(g)
    ref = odist.run(meas, g.num_poses, 4, opts, X0, 8, "hash", log_global=False)
    got = cpu_ref.run(meas, g.num_poses, 4, opts, X0, 8, "hash", workers=1, threads=2)
    a, b = np.array(ref["fobj_nodes"]), np.array(got["fobj_nodes"])
    assert np.abs(a - b).max() <= 1e-9 * np.abs(a).max()


def test_cpu_baseline_larger_nodes_sparse_cholesky():
    """1 600-pose nodes: the up-looking sparse Cholesky with the supplied ordering on a non-trivial pattern."""
    g, _, X0 = D.grid3d(20, 20, 12, seed=8)
    ref, got = _both(g, 3, X0, 4, "star", "trivial")
    a, b = np.array(ref["fobj_nodes"]), np.array(got["fobj_nodes"])
    assert np.abs(a - b).max() <= 1e-9 * np.abs(a).max()
    assert np.abs(ref["X"] - got["X"]).max() < 1e-7
