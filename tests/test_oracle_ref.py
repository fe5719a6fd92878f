"""Pins the oracle against REAL reference code compiled in place (oracle/_ref, built by
oracle/Makefile from /root/reference): the AVX2 SO(2)/SO(3) polar projections
(C++/DPGO/src/internal/project_to_SOd.cpp) and the header-only TNT / STPCG
(C++/Optimization/include/Optimization/Riemannian/TNT.h, LinearAlgebra/IterativeSolvers.h)."""
import math

import numpy as np
import pytest

import dpgo_b200.graph as G
import parity
from oracle import dist_pgo as odist
from oracle import dpgo as odpgo
from oracle import ref, sod, solver

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def _inputs(d, rng):
    n = 515                                     # not a multiple of 4: exercises the tail group
    A = rng.standard_normal((n, d, d))
    A[100:200] *= 1e-6                          # tiny
    A[200:300] *= 1e6                           # huge
    Q = sod.project_svd(rng.standard_normal((100, d, d)))
    A[300:400] = Q + 1e-4 * rng.standard_normal((100, d, d))      # near rotations (the usual case)
    S = np.zeros((50, d, d))
    for k in range(d):
        S[:, k, k] = d - k
    S[:, d - 1, d - 1] = -1.0                                     # negative determinant, distinct singular values
    A[400:450] = Q[:50] @ S @ np.swapaxes(Q[50:100], 1, 2)
    A[450:470, :, 0] = A[450:470, :, 1]                           # rank deficient
    return A


@pytest.mark.parametrize("d", [2, 3])
def test_projection_matches_reference_avx2(d):
    rng = np.random.default_rng(7)
    A = _inputs(d, rng)
    want = ref.project(A, nofma=True)
    got = sod.project_to_SO3(A) if d == 3 else sod.project_to_SO2(A)
    scale = 1.0
    # numpy evaluates the reference's fma(a,b,c) as a*b+c: <= a few ulp per entry
    ok = np.abs(got - want) <= 2e-14 * scale
    bad = ~ok.all(axis=(1, 2))
    # the rank-deficient block may legitimately differ (the projection is not unique there)
    assert not bad[:450].any(), np.abs(got - want)[:450].max()
    # GCC contracts the reference's mul/add pairs into FMAs in the stock build: same answer to ulps
    assert np.abs(ref.project(A) - want)[:450].max() < 2e-14
    # and both are the polar factor
    U = want[:450]
    assert np.abs(U @ np.swapaxes(U, 1, 2) - np.eye(d)).max() < 1e-12
    assert np.abs(np.linalg.det(U) - 1).max() < 1e-12


def _both_tnt(f, QM, metric, retract, x0, precon, params):
    """Drop-in for oracle.solver.tnt that also runs the reference's TNT.h on the same callbacks."""
    a = solver.tnt(f, QM, metric, retract, x0, precon, params)
    b = ref.tnt(f, QM, metric, retract, x0, precon, params)
    _both_tnt.calls += 1
    assert a.inner_iterations == b.inner_iterations, (a.inner_iterations, b.inner_iterations)
    assert np.allclose(a.gain_ratios, b.gain_ratios, rtol=1e-7, atol=1e-9, equal_nan=True)
    assert abs(a.f - b.f) <= 1e-12 * max(1.0, abs(a.f))
    assert np.abs(a.x - b.x).max() <= 1e-11
    return a


@pytest.mark.parametrize("loss,precon", [("trivial", "BlockJacobi"), ("trivial", "None"), ("huber", "Jacobi")])
def test_tnt_and_stpcg_match_reference_headers(monkeypatch, loss, precon):
    # every truncated-Newton call of a short AMM-PGO# run goes through both implementations
    g, _, X0 = G.grid3d(5, 5, 4, seed=3)
    _both_tnt.calls = 0
    monkeypatch.setattr(odpgo, "tnt", _both_tnt)
    opts = odpgo.Options(loss=loss, preconditioner=precon)
    odist.run(parity.to_measurements(g), g.num_poses, 4, opts, X0, 5, "hash")
    assert _both_tnt.calls >= 8


def test_tnt_sphere_known_answer_in_both():
    # C++/Optimization/tests/TNT_unit_test.cpp:63-187: minimise |X - P|^2 over S^2
    P = np.array([0.0, 0.0, 1.0])
    X0 = np.array([-0.5, -0.5, -0.707107])
    proj = lambda X, V: V - X.dot(V) * X
    f = lambda X: float((X - P) @ (X - P))
    grad = lambda X: proj(X, 2 * (X - P))
    QM = lambda X: (grad(X), lambda X_, V: proj(X_, 2 * V) - X_.dot(grad(X_)) * V)
    metric = lambda X, a, b: float(a @ b)
    retract = lambda X, V: (X + V) / np.linalg.norm(X + V)
    for precon in (None, lambda X, V: np.array([1.0, 2.0, 3.0]) * V):
        params = solver.TNTParams()
        params.relative_decrease_tolerance = 0
        params.stepsize_tolerance = 0
        params.preconditioned_gradient_tolerance = 0
        params.gradient_tolerance = 1e-8
        a = solver.tnt(f, QM, metric, retract, X0, precon, params)
        b = ref.tnt(f, QM, metric, retract, X0, precon, params)
        assert a.inner_iterations == b.inner_iterations
        assert np.abs(a.x - b.x).max() < 1e-12 and np.abs(a.x - P).max() < 1e-7
        assert math.sqrt(metric(a.x, grad(a.x), grad(a.x))) < 1e-8
