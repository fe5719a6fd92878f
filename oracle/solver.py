"""Truncated-Newton trust region + Steihaug-Toint PCG (oracle; test
infrastructure only).

Restates C++/Optimization/include/Optimization/LinearAlgebra/
IterativeSolvers.h:166-426 (STPCG, without the constraint-preconditioner
branches `At`, which DPGO never passes) and
C++/Optimization/include/Optimization/Riemannian/TNT.h:242-693.
"""
from __future__ import annotations

import math

import numpy as np


def stpcg(g, H, inner, Delta, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
          P=None, epsilon=1e-8, Hs_out=None):
    """Returns (s, update_step_M_norm, num_iterations).
    IterativeSolvers.h:166-426.

    Hs_out (a list, test instrumentation): receives H s accumulated from the
    products H p_k the iteration forms anyway (sum_k alpha_k H p_k, plus
    sigma H p_k on a boundary exit) -- what the CUDA path feeds the model
    decrease with instead of the reference's extra Hess(x, h) of TNT.h:514-515."""
    if Delta <= 0:
        raise ValueError("Trust-region radius (Delta) must be a positive real value")
    if kappa_fgr < 0 or kappa_fgr >= 1:
        raise ValueError("kappa_fgr must be in [0,1)")
    if theta < 0 or theta > 1:
        raise ValueError("theta must be in [0,1]")
    if epsilon <= 0 or epsilon >= 1:
        raise ValueError("epsilon must be in (0,1)")
    s_k = 0 * g
    r_k = g.copy()
    v_k = r_k if P is None else P(r_k)
    p_k = -v_k
    sk_M_pk = 0.0
    sk_M_2 = 0.0
    pk_M_2 = inner(r_k, v_k)
    Delta_2 = Delta * Delta
    r0_norm = math.sqrt(inner(r_k, v_k))
    target = r0_norm * min(kappa_fgr, math.pow(r0_norm, theta))
    num_iterations = 0
    Hs = 0 * g
    if Hs_out is not None:
        Hs_out.append(Hs)
    while num_iterations < max_iterations:
        if math.sqrt(inner(r_k, v_k)) <= target:
            break
        Hp_k = H(p_k)
        kappa_k = inner(p_k, Hp_k)
        if math.sqrt(inner(Hp_k, Hp_k)) / math.sqrt(inner(p_k, p_k)) < epsilon:
            if inner(p_k, r_k) < 0:
                p_k = -p_k
                Hp_k = -Hp_k
                sk_M_pk = -sk_M_pk
            sigma = (-sk_M_pk + math.sqrt(sk_M_pk * sk_M_pk +
                                          pk_M_2 * (Delta_2 - sk_M_2))) / pk_M_2
            if Hs_out is not None:
                Hs_out[0] = Hs + sigma * Hp_k
            return s_k + sigma * p_k, Delta, num_iterations
        alpha_k = inner(r_k, v_k) / kappa_k
        skp1_M_2 = sk_M_2 + 2 * alpha_k * sk_M_pk + alpha_k * alpha_k * pk_M_2
        if kappa_k <= 0 or skp1_M_2 > Delta_2:
            sigma = (-sk_M_pk + math.sqrt(sk_M_pk * sk_M_pk +
                                          pk_M_2 * (Delta_2 - sk_M_2))) / pk_M_2
            if Hs_out is not None:
                Hs_out[0] = Hs + sigma * Hp_k
            return s_k + sigma * p_k, Delta, num_iterations
        s_k = s_k + alpha_k * p_k
        Hs = Hs + alpha_k * Hp_k
        if Hs_out is not None:
            Hs_out[0] = Hs
        r_k = r_k + alpha_k * Hp_k
        v_k = r_k if P is None else P(r_k)
        rk_vk = inner(r_k, v_k)
        beta_k = rk_vk / (alpha_k * kappa_k)
        sk_M_2 = skp1_M_2
        sk_M_pk = beta_k * (sk_M_pk + alpha_k * pk_M_2)
        pk_M_2 = rk_vk + beta_k * beta_k * pk_M_2
        p_k = -v_k + beta_k * p_k
        num_iterations += 1
    return s_k, math.sqrt(sk_M_2), num_iterations


class TNTParams:
    """TNT.h:76-129 defaults + SmoothOptimizerParams."""
    def __init__(self):
        self.gradient_tolerance = 1e-6
        self.relative_decrease_tolerance = 1e-6
        self.stepsize_tolerance = 1e-6
        self.max_iterations = 1000
        self.max_iterations_accepted = 10 ** 9
        self.Delta0 = 1.0
        self.eta1 = 0.05
        self.eta2 = 0.9
        self.alpha1 = 0.25
        self.alpha2 = 2.5
        self.max_TPCG_iterations = 1000
        self.kappa_fgr = 0.1
        self.theta = 0.5
        self.preconditioned_gradient_tolerance = 1e-6
        self.Delta_tolerance = 1e-6


class TNTResult:
    pass


def tnt(f, QM, metric, retract, x0, precon, params, accumulated_Hs=False):
    """TNT.h:242-693.  QM(x) -> (grad, Hess(x, v)); metric(x, v1, v2);
    retract(x, h); precon(x, v) or None.

    accumulated_Hs=True (test instrumentation) evaluates the model decrease with
    H h accumulated inside STPCG, as the CUDA path does, instead of the fresh
    Hess(x, h) of TNT.h:514-515; res.gain_ratios_fresh then holds the
    reference's value of every gain ratio next to the one that was used."""
    sqrt_eps = math.sqrt(np.finfo(float).eps)
    res = TNTResult()
    res.status = "IterationLimit"
    res.inner_iterations = []
    res.gain_ratios = []
    res.gain_ratios_fresh = []
    x = x0.copy()
    fx = f(x)
    grad, Hess = QM(x)
    gnorm = math.sqrt(metric(x, grad, grad))
    if precon is not None:
        pg = precon(x, grad)
        pgnorm = math.sqrt(metric(x, pg, pg))
    else:
        pgnorm = gnorm
    Delta = params.Delta0
    it = 0
    acc = 0
    while it < params.max_iterations and acc < params.max_iterations_accepted:
        if gnorm < params.gradient_tolerance:
            res.status = "Gradient"
            break
        if pgnorm < params.preconditioned_gradient_tolerance:
            res.status = "PreconditionedGradient"
            break
        Hx = (lambda v, x=x, Hess=Hess: Hess(x, v))
        ip = (lambda a, b, x=x: metric(x, a, b))
        Px = None if precon is None else (lambda v, x=x: precon(x, v))
        Hs_acc = []
        h, h_M_norm, inner_its = stpcg(grad, Hx, ip, Delta,
                                       params.max_TPCG_iterations,
                                       params.kappa_fgr, params.theta, Px,
                                       Hs_out=Hs_acc)
        h_norm = math.sqrt(metric(x, h, h))
        x_prop = retract(x, h)
        fx_prop = f(x_prop)
        dm = -metric(x, grad, h) - 0.5 * metric(x, h, Hess(x, h))
        df = fx - fx_prop
        rel_dec = df / (sqrt_eps + abs(fx))
        with np.errstate(all="ignore"):
            rho = float(np.float64(df) / np.float64(dm))
        if accumulated_Hs:
            res.gain_ratios_fresh.append(rho)
            dm = -metric(x, grad, h) - 0.5 * metric(x, h, Hs_acc[0])
            with np.errstate(all="ignore"):
                rho = float(np.float64(df) / np.float64(dm))
        accepted = (not math.isnan(rho)) and rho > params.eta1
        acc += int(accepted)
        res.inner_iterations.append(inner_its)
        res.gain_ratios.append(rho)
        stop = False
        if accepted:
            x = x_prop
            fx = fx_prop
            if rel_dec < params.relative_decrease_tolerance:
                res.status = "RelativeDecrease"
                break
            if h_norm < params.stepsize_tolerance:
                res.status = "Stepsize"
                break
            grad, Hess = QM(x)
            gnorm = math.sqrt(metric(x, grad, grad))
            if precon is not None:
                pg = precon(x, grad)
                pgnorm = math.sqrt(metric(x, pg, pg))
            else:
                pgnorm = gnorm
        if (not math.isnan(rho)) and rho >= params.eta2:
            Delta = max(params.alpha2 * h_M_norm, Delta)
        elif math.isnan(rho) or rho < params.eta1:
            Delta = params.alpha1 * h_M_norm
            if Delta < params.Delta_tolerance:
                res.status = "TrustRegion"
                break
        it += 1
    res.x = x
    res.f = fx
    res.gradfx_norm = gnorm
    res.iterations = it
    return res
