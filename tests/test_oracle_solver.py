"""Pins the oracle's STPCG / TNT against the reference's own known-answer tests:
C++/Optimization/tests/IterativeSolvers_unit_test.cpp:140-330 and
C++/Optimization/tests/TNT_unit_test.cpp:63-187 (same matrices, start points,
tolerances and expectations)."""
import numpy as np

from oracle.solver import TNTParams, stpcg, tnt

EPS_ABS = 1e-6
EPS_REL = 1e-6
small_g = np.array([21.0, -0.4, 19.0])
small_P = np.array([1000.0, 100.0, 1.0])
small_M = np.array([100.0, 10.0, 1.0])
FMAX = np.finfo(float).max
inner = lambda a, b: float(a @ b)


def test_exact_stpcg():
    s, nrm, _ = stpcg(small_g, lambda v: small_P * v, inner, FMAX, 3, 1e-8, 0.999)
    assert np.linalg.norm(s + small_g / small_P) < EPS_ABS
    assert abs(nrm - np.linalg.norm(s)) / np.linalg.norm(s) < EPS_REL


def test_exact_stpcg_negative_curvature():
    Delta = 1000.0
    s, nrm, _ = stpcg(small_g, lambda v: -small_P * v, inner, Delta, 3, 1e-8, 0.999)
    assert np.linalg.norm(s + Delta / np.linalg.norm(small_g) * small_g) < EPS_ABS
    assert abs(nrm - np.linalg.norm(s)) / np.linalg.norm(s) < EPS_REL


def test_exact_stpcg_preconditioned():
    s, nrm, _ = stpcg(small_g, lambda v: small_P * v, inner, FMAX, 3, 1e-8, 0.999,
                      P=lambda v: v / small_M)
    assert np.linalg.norm(s + small_g / small_P) < EPS_ABS
    sM = np.sqrt(s @ (small_M * s))
    assert abs((nrm - sM) / sM) < EPS_REL


def test_exact_stpcg_negative_curvature_preconditioned():
    Delta = 1000.0
    s, nrm, _ = stpcg(small_g, lambda v: -small_P * v, inner, Delta, 3, 1e-8, 0.999,
                      P=lambda v: v / small_M)
    p = -small_g / small_M
    s_gt = Delta / np.sqrt(p @ (small_M * p)) * p
    assert np.linalg.norm(s - s_gt) < EPS_ABS
    sM = np.sqrt(s @ (small_M * s))
    assert abs(nrm - sM) / sM < EPS_REL


def test_stpcg_truncation():
    rng = np.random.default_rng(0)
    n = 1000
    g = rng.uniform(-1, 1, n)
    P = 2000 + 1000 * rng.uniform(-1, 1, n)
    s, nrm, _ = stpcg(g, lambda v: P * v, inner, 1000.0, 3, 0.1, 0.7)
    assert np.linalg.norm(g + P * s) / np.linalg.norm(g) < 0.1
    assert abs(nrm - np.linalg.norm(s)) / np.linalg.norm(s) < EPS_REL


def test_stpcg_preconditioned_truncation():
    rng = np.random.default_rng(1)
    n = 1000
    g = rng.uniform(-1, 1, n)
    P = 2000 + 1000 * rng.uniform(-1, 1, n)
    M = 2000 + 1000 * rng.uniform(-1, 1, n)
    s, nrm, _ = stpcg(g, lambda v: P * v, inner, 1000.0, n, 0.1, 0.7, P=lambda v: v / M)
    r = g + P * s
    assert np.sqrt(r @ (r / M)) / np.sqrt(g @ (g / M)) < 0.1
    sM = np.sqrt(s @ (M * s))
    assert abs((nrm - sM) / sM) < EPS_REL


def _sphere_problem():
    Pn = np.array([0.0, 0.0, 1.0])
    project = lambda X, V: V - (X @ V) * X
    F = lambda X: float(np.sum((X - Pn) ** 2))
    gradF = lambda X: project(X, 2 * (X - Pn))

    def QM(X):
        return gradF(X), (lambda X_, Xdot: project(X_, 2 * Xdot) - (X_ @ gradF(X_)) * Xdot)
    metric = lambda X, a, b: float(a @ b)
    retract = lambda X, V: (X + V) / np.linalg.norm(X + V)
    X0 = np.array([-0.5, -0.5, -0.707107])
    return F, gradF, QM, metric, retract, X0


def _params():
    p = TNTParams()
    p.relative_decrease_tolerance = 0
    p.stepsize_tolerance = 0
    p.preconditioned_gradient_tolerance = 0
    p.gradient_tolerance = 1e-8
    return p


def test_tnt_sphere():
    F, gradF, QM, metric, retract, X0 = _sphere_problem()
    res = tnt(F, QM, metric, retract, X0, None, _params())
    assert res.status == "Gradient"
    assert np.linalg.norm(gradF(res.x)) < 1e-8
    assert F(res.x) < F(X0)


def test_tnt_sphere_with_precon():
    F, gradF, QM, metric, retract, X0 = _sphere_problem()
    precon = lambda X, V: np.array([1.0, 2.0, 3.0]) * V
    res = tnt(F, QM, metric, retract, X0, precon, _params())
    assert res.status == "Gradient"
    assert np.linalg.norm(gradF(res.x)) < 1e-8
    assert F(res.x) < F(X0)
