"""DPGOProblem / DPGOHash / DPGOStar (oracle; test infrastructure only).

Restates C++/DPGO/src/DPGOProblem.cpp, C++/DPGO/src/DPGOHash.cpp and
C++/DPGO/src/DPGOStar.cpp with the Static rescale path (the one `dist_pgo`
uses, C++/examples/dist_pgo.cpp:105).  Matrices use the reference layout:
X is ((d+1) n0, d) = [t rows; R rows], Z additionally carries the neighbour
copies [t_nbr; R_nbr] (C++/DPGO/src/DPGO_utils.cpp:413-424).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import data_matrix as dm
from . import sod
from .solver import TNTParams, tnt


class Options:
    """DPGO::Options with the values dist_pgo sets (dist_pgo.cpp:103-120 over
    DPGO_types.h:78-201)."""

    def __init__(self, **kw):
        self.scheme = "AMM"
        self.regularizer = 1e-11
        self.accepted_delta = 5e-4
        self.eta = (5e-4, 2.5e-2)
        self.psi = 1e-10
        self.phi = 1e-6
        self.max_soft_restart_hits = (10, 25)
        self.oscillation_cnt_period = 15
        self.max_oscillations = 12
        self.loss = "trivial"          # trivial | huber | gm | welsch
        self.loss_reg = 0.25
        self.rescale = "Static"        # Static (what dist_pgo sets, dist_pgo.cpp:105) | Dynamic (DPGO_types.h:128)
        self.max_rescale_count = 5     # DPGO_types.h:131
        self.grad_norm_tol = 1e-3
        self.rel_func_decrease_tol = 1e-6
        self.stepsize_tol = 1e-4
        self.max_iterations = 10
        self.max_iterations_accepted = 1
        self.preconditioner = "RegularizedCholesky"
        self.reg_Cholesky_precon_max_condition_number = 1e6
        self.preconditioned_grad_norm_tol = 1e-4
        self.max_tCG_iterations = 10000
        self.STPCG_kappa = 0.05
        self.STPCG_theta = 0.9
        self.lambda_max_override = None   # share Spectra's loose estimate
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


def loss_weights(e, loss, delta):
    """IRLS weights and loss values of DPGOProblem::evaluate_E
    (DPGOProblem.cpp:647-675).  Returns (omega, fobjE)."""
    if loss == "trivial":
        return np.ones_like(e), 0.5 * e.sum()
    if loss == "huber":
        resc = np.sqrt(np.maximum(e, delta))
        w = math.sqrt(delta) / resc
        return w, 0.5 * np.minimum(2 * math.sqrt(delta) * resc - delta, e).sum()
    if loss == "gm":
        w = (delta * delta) / np.square(e + delta)
        return w, 0.5 * delta * (e / (e + delta)).sum()
    if loss == "welsch":
        w = np.exp(-e / delta)
        return w, 0.5 * (delta * len(e) - delta * w.sum())
    raise ValueError(loss)


def _tr(A, B):
    return float(np.sum(A * B))


class DPGOProblem:
    """DPGOProblem.cpp:11-125 (constructor) and the operators below it."""

    def __init__(self, node, meas, opts):
        self.node = node
        self.opts = opts
        self.loss = opts.loss
        self.info = info = dm.generate_data_info(node, meas)
        self.d = d = info.d
        self.n = info.n
        self.m = info.m
        self.quadratic = opts.loss == "trivial"     # SIMPLE == 1
        self.dynamic = (opts.rescale == "Dynamic") and not self.quadratic
        self.DiagReScale = np.ones(self.m[1])                    # DPGOProblem.cpp:31
        mats = dm.build_data_matrices(info, opts.regularizer, self.quadratic,
                                      self.DiagReScale if self.dynamic else None, self.dynamic)
        self.__dict__.update(mats)
        n0 = self.n[0]
        self.size0 = (d + 1) * n0
        self.L = spla.splu(sp.csc_matrix(self.G00),
                           permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                           options=dict(SymmetricMode=True))
        self.precon_kind = opts.preconditioner
        if self.precon_kind == "Jacobi":            # :96-98
            self.jacobi = 1.0 / self.G11.diagonal()
        elif self.precon_kind == "BlockJacobi":
            # d x d diagonal blocks of G11 (not in the reference; the
            # north-star preconditioner of the CUDA path)
            G11 = self.G11.tocsr()
            blocks = np.zeros((n0, d, d))
            coo = G11.tocoo()
            sel = (coo.row // d) == (coo.col // d)
            np.add.at(blocks, (coo.row[sel] // d, coo.row[sel] % d,
                               coo.col[sel] % d), coo.data[sel])
            self.block_jacobi = np.linalg.inv(blocks)
        elif self.precon_kind == "RegularizedCholesky":   # :101-124
            if opts.lambda_max_override is not None:
                lam = float(opts.lambda_max_override[node]) \
                    if hasattr(opts.lambda_max_override, "__len__") \
                    else float(opts.lambda_max_override)
            else:
                if self.G11.shape[0] > 3:
                    lam = float(spla.eigsh(self.G11, k=1, which="LM", tol=1e-4,
                                           return_eigenvectors=False)[0])
                else:
                    lam = float(np.linalg.eigvalsh(self.G11.toarray())[-1])
            self.lambda_max = lam
            A = self.G11 + sp.identity(self.G11.shape[0]) * (
                lam / opts.reg_Cholesky_precon_max_condition_number)
            self.reg_chol = spla.splu(sp.csc_matrix(A),
                                      permc_spec="MMD_AT_PLUS_A",
                                      diag_pivot_thresh=0.0,
                                      options=dict(SymmetricMode=True))
        elif self.precon_kind != "None":
            raise ValueError(self.precon_kind)

    # ---- Rescale::Dynamic -------------------------------------------------
    MAX_RESCALE, MIN_RESCALE = 1.0, 0.01                         # DPGOProblem.h:17-18

    def update_quadratic_mat(self, DiagReScale):
        """DPGOProblem.cpp:751-840 + L_.factorize (:315, :479): the majoriser G, D, Q and the auxiliary T, N, V at
        the new rescale vector.  (The preconditioner is NOT refreshed, as in the reference.)"""
        self.DiagReScale = DiagReScale
        mats = dm.build_data_matrices(self.info, self.opts.regularizer, False, DiagReScale, True)
        for k in ("G", "D", "Q", "H", "G00", "G01", "G10", "G11", "T", "N", "V"):
            setattr(self, k, mats[k])
        self.L = spla.splu(sp.csc_matrix(self.G00), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                           options=dict(SymmetricMode=True))

    def _maybe_rescale(self, w, rescale_count, max_rescale_count):
        """DPGOProblem.cpp:301-321 / :465-485.  Returns the new rescale_count."""
        rescaled = (rescale_count >= max_rescale_count) or bool(np.any(w > self.DiagReScale))
        if rescaled:
            self.update_quadratic_mat(np.clip(1.25 * w, self.MIN_RESCALE, self.MAX_RESCALE))
            return 0
        return rescale_count + 1

    def evaluate_g_and_f0_rescale(self, Z, rescale_count, max_rescale_count):
        """DPGOProblem.cpp:289-358 -> (g, f0, Dfobj, fobj, DfobjE, fobjE, rescale_count)."""
        s0 = self.size0
        w, DfobjE, fobjE = self.evaluate_E(Z)
        if self.dynamic:
            rescale_count = self._maybe_rescale(w, rescale_count, max_rescale_count)
        X = Z[:s0]
        g = DfobjE[:s0].copy()
        temp = self.D @ X
        g -= temp
        temp = 0.5 * temp - DfobjE[:s0]
        f0 = 0.5 * fobjE + _tr(X, temp)
        temp = self.G @ X
        Dfobj = g + temp
        temp = 0.5 * temp + g
        fobj = f0 + _tr(X, temp)
        return g, f0, Dfobj, fobj, DfobjE, fobjE, rescale_count

    def evaluate_g_and_f_rescale(self, Z, Z0, G, DfobjE0, fobjE0, rescale_count, max_rescale_count):
        """DPGOProblem.cpp:426-514 (robust branch) -> (g, f, Dfobj, fobj, DfobjE, fobjE, rescale_count).  The
        history term uses Q BEFORE the rescale, g / Dfobj / f use D and G AFTER it."""
        s0 = self.size0
        X = Z[:s0]
        Y = Z - Z0
        temp = DfobjE0 + 0.5 * (self.Q @ Y)
        fobj = G - 0.5 * fobjE0
        fobj -= 0.5 * _tr(Y, temp)
        w, DfobjE, fobjE = self.evaluate_E(Z)
        fobj += 0.5 * fobjE
        if self.dynamic:
            rescale_count = self._maybe_rescale(w, rescale_count, max_rescale_count)
        g = DfobjE[:s0] - self.D @ X
        temp = self.G @ X
        Dfobj = g + temp
        temp = 0.5 * temp + g
        f = fobj - _tr(X, temp)
        return g, f, Dfobj, fobj, DfobjE, fobjE, rescale_count

    # ---- geometry -------------------------------------------------------
    def recover_translations(self, R, g):
        """DPGOProblem.h:275-294."""
        n0 = self.n[0]
        temp = g[:n0] + self.G01 @ R
        return -self.L.solve(temp)

    def retract(self, Y, Ydot, g):
        """DPGOProblem.cpp:127-143."""
        n0, d = self.n[0], self.d
        Rp = sod.project(Y[n0:] + Ydot, d)
        tp = self.recover_translations(Rp, g)
        return np.vstack([tp, Rp])

    def full_tangent_space_projection(self, Y, Ydot):
        """DPGOProblem.cpp:145-162."""
        n0 = self.n[0]
        return np.vstack([Ydot[:n0], sod.proj(Y[n0:], Ydot[n0:], self.d)])

    def reduced_tangent_space_projection(self, Y, Ydot):
        """DPGOProblem.cpp:164-178."""
        return sod.proj(Y[self.n[0]:], Ydot, self.d)

    # ---- surrogate G(X|Z) -----------------------------------------------
    def evaluate_G(self, Y, g, f):
        """DPGOProblem.cpp:180-203: tr(Y^T (g + 1/2 G Y)) + f."""
        temp = g + 0.5 * (self.G @ Y)
        return _tr(Y, temp) + f

    def evaluate_E(self, Z):
        """DPGOProblem.cpp:634-681.  Returns (omega, DfobjE, fobjE)."""
        d, m1 = self.d, self.m[1]
        Err = self.B1 @ Z                        # ((d+1) m1, d)
        e = np.square(Err).reshape(m1, (d + 1) * d).sum(axis=1)
        w, fobjE = loss_weights(e, self.loss, self.opts.loss_reg)
        Wrow = np.repeat(w, d + 1)[:, None]
        DfobjE = self.B1.T @ (Wrow * Err)
        self.last_weights = w
        self.last_sq_err = e
        return w, DfobjE, float(fobjE)

    def evaluate_g_and_f0(self, Z):
        """DPGOProblem.cpp:222-267 -> (g, f0, Dfobj, fobj, DfobjE, fobjE)."""
        s0 = self.size0
        _, DfobjE, fobjE = self.evaluate_E(Z)
        X = Z[:s0]
        g = DfobjE[:s0].copy()
        temp = self.D @ X
        g -= temp
        temp = 0.5 * temp - DfobjE[:s0]
        f0 = 0.5 * fobjE + _tr(X, temp)
        temp = self.G @ X
        Dfobj = g + temp
        temp = 0.5 * temp + g
        fobj = f0 + _tr(X, temp)
        return g, f0, Dfobj, fobj, DfobjE, fobjE

    def evaluate_none_g_and_f0(self, Z):
        """DPGOProblem.cpp:269-287."""
        g = self.S @ Z
        f0 = 0.5 * _tr(Z, self.P0 @ Z)
        return g, f0

    def evaluate_g_and_f(self, Z, Z0, G, DfobjE0, fobjE0):
        """DPGOProblem.cpp:360-424 (robust branch) ->
        (g, f, Dfobj, fobj, DfobjE, fobjE)."""
        s0 = self.size0
        X = Z[:s0]
        Y = Z - Z0
        temp = DfobjE0 + 0.5 * (self.Q @ Y)
        fobj = G - 0.5 * fobjE0
        fobj -= 0.5 * _tr(Y, temp)
        _, DfobjE, fobjE = self.evaluate_E(Z)
        fobj += 0.5 * fobjE
        g = DfobjE[:s0] - self.D @ X
        temp = self.G @ X
        Dfobj = g + temp
        temp = 0.5 * temp + g
        f = fobj - _tr(X, temp)
        return g, f, Dfobj, fobj, DfobjE, fobjE

    def evaluate_none_g_and_f(self, Z, Z0, G):
        """DPGOProblem.cpp:516-542 -> (g, f, fobj)."""
        g = self.S @ Z
        Y = Z - Z0
        fobj = G + 0.5 * _tr(Y, self.Q @ Y)
        f = fobj + 0.5 * _tr(Z, self.P @ Z)
        return g, f, fobj

    def evaluate_g(self, Z):
        """DPGOProblem.cpp:683-725."""
        if self.quadratic:
            return self.S @ Z
        s0 = self.size0
        _, DfobjE, _ = self.evaluate_E(Z)
        return DfobjE[:s0] - self.D @ Z[:s0]

    def evaluate_g_and_Df(self, Z):
        """DPGOProblem.cpp:544-550, :727-749."""
        g = self.evaluate_g(Z)
        return g, g + self.G @ Z[:self.size0]

    # ---- inner-solve operators -------------------------------------------
    def reduced_Euclidean_gradient_G(self, Y, g):
        """DPGOProblem.h:370-382."""
        n0 = self.n[0]
        return g[n0:] + (self.G[n0:] @ Y)

    def reduced_Hess_vec(self, Y, nablaF_Y, Ydot):
        """DPGOProblem.cpp:552-577."""
        n0, d = self.n[0], self.d
        R = Y[n0:]
        tdot = -self.L.solve(self.G01 @ Ydot)
        E = self.G10 @ tdot + self.G11 @ Ydot
        E -= sod.sym_block_diag_product(Ydot, R, nablaF_Y, d)
        return sod.proj(R, E, d)

    def precondition(self, Y, Ydot):
        """DPGOProblem.cpp:579-598."""
        k = self.precon_kind
        if k == "Jacobi":
            return self.reduced_tangent_space_projection(Y, self.jacobi[:, None] * Ydot)
        if k == "BlockJacobi":
            d = self.d
            v = (self.block_jacobi @ Ydot.reshape(-1, d, d)).reshape(-1, d)
            return self.reduced_tangent_space_projection(Y, v)
        if k == "RegularizedCholesky":
            return self.reduced_tangent_space_projection(Y, self.reg_chol.solve(Ydot))
        return Ydot

    def proximal(self, Z, Df):
        """DPGOProblem.cpp:600-632."""
        n0, d = self.n[0], self.d
        t0 = Z[:n0]
        R0 = Z[n0:(d + 1) * n0]
        if self.quadratic:
            M = self.U @ Z
        else:
            M = -Df[n0:] + self.N.T @ Df[:n0] + self.V @ R0
        R = sod.project(M, d)
        t = t0 - self.N @ (R - R0) - self.T[:, None] * Df[:n0]
        return np.vstack([t, R])


class NodeState:
    """The fields of DPGOResult (DPGO_types.h:204-322) the drivers use; only
    iterates k and k-1 are kept."""

    def __init__(self):
        self.updated = True
        self.iters = 0
        self.soft_restart_hits = [0, 0]
        self.oscillations = []
        self.num_oscillations = 0
        self.gamma = 0.0
        self.s = {}
        self.Fk = [0.0, 0.0]
        self.Gk = 0.0
        self.fobj_prev = None
        self.X_prev = None
        self.g_prev = None
        self.Dfobj_prev = None
        self.tcg_iters = 0
        self.rescale_count = 0           # DPGO_types.h:277
        self.n_restarts = 0
        self.last_refined = False


class _NodeOps:
    """The five TNT callbacks built in DPGOHash::amm_pgo (DPGOHash.cpp:270-331)
    and the parameter block (:337-349)."""

    def __init__(self, problem, opts):
        self.problem = problem
        self.opts = opts

    def run_tnt(self, x0, g, f, st):
        p, o = self.problem, self.opts
        params = TNTParams()
        params.gradient_tolerance = o.grad_norm_tol
        params.preconditioned_gradient_tolerance = o.preconditioned_grad_norm_tol
        params.relative_decrease_tolerance = o.rel_func_decrease_tol
        params.stepsize_tolerance = o.stepsize_tol
        params.max_iterations = o.max_iterations
        params.max_iterations_accepted = o.max_iterations_accepted
        params.max_TPCG_iterations = o.max_tCG_iterations
        params.kappa_fgr = o.STPCG_kappa
        params.theta = o.STPCG_theta
        cache = {}

        def fobj(Y):
            return p.evaluate_G(Y, g, f)

        def QM(Y):
            nab = p.reduced_Euclidean_gradient_G(Y, g)
            grad = p.reduced_tangent_space_projection(Y, nab)
            cache["nab"] = nab
            return grad, (lambda Y_, v, nab=nab: p.reduced_Hess_vec(Y_, nab, v))

        def metric(Y, a, b):
            return _tr(a, b)

        def retract(Y, h):
            return p.retract(Y, h, g)

        precon = None if p.precon_kind == "None" else \
            (lambda Y, v: p.precondition(Y, v))
        res = tnt(fobj, QM, metric, retract, x0, precon, params)
        st.tcg_iters += sum(res.inner_iterations)
        return res


class DPGOHash:
    """AMM-PGO# / MM-PGO per-node driver (DPGOHash.cpp)."""

    def __init__(self, node, meas, opts):
        self.opts = opts
        self.problem = DPGOProblem(node, meas, opts)
        self.ops = _NodeOps(self.problem, opts)
        self.st = NodeState()

    # DPGOHash.cpp:20-43
    def initialize(self, X):
        p = self.problem
        assert X.shape == ((p.d + 1) * (p.n[0] + p.n[1]), p.d)
        st = self.st = NodeState()
        st.Xk = X.copy()
        st.Xak = st.Xk[:p.size0].copy()
        st.Xakh = np.zeros_like(st.Xak)
        st.gamma = 0.0
        st.updated = False
        return 0

    # DPGOHash.cpp:84-228
    def update(self):
        st, p, o = self.st, self.problem, self.opts
        if st.updated:
            return 0
        it = st.iters
        X_it = st.Xk.copy()
        if p.quadratic:
            if it == 0:
                g, f = p.evaluate_none_g_and_f0(X_it)
                fobj = p.evaluate_G(st.Xak, g, f)
            else:
                g, f, fobj = p.evaluate_none_g_and_f(X_it, st.X_cur, st.Gk)
            Dfobj = None
        elif o.rescale == "Static":
            if it == 0:
                g, f, Dfobj, fobj, st.DfobjE, st.fobjE = p.evaluate_g_and_f0(X_it)
            else:
                g, f, Dfobj, fobj, st.DfobjE, st.fobjE = p.evaluate_g_and_f(
                    X_it, st.X_cur, st.Gk, st.DfobjE, st.fobjE)
        else:                                                   # DPGOHash.cpp:131-143
            if it == 0:
                g, f, Dfobj, fobj, st.DfobjE, st.fobjE, st.rescale_count = p.evaluate_g_and_f0_rescale(
                    X_it, st.rescale_count, o.max_rescale_count)
            else:
                g, f, Dfobj, fobj, st.DfobjE, st.fobjE, st.rescale_count = p.evaluate_g_and_f_rescale(
                    X_it, st.X_cur, st.Gk, st.DfobjE, st.fobjE, st.rescale_count, o.max_rescale_count)
        if it == 0:
            st.Fk = [fobj, fobj]
            st.Gk = fobj
        if p.quadratic:
            # full_Riemannian_gradient_G, DPGOProblem.h:356-368,398-402
            Dfobj = g + p.G @ st.Xak
        st.gradF = p.full_tangent_space_projection(st.Xak, Dfobj)
        st.gradFnorm = float(np.linalg.norm(st.gradF))
        # shift history (reference keeps every iterate; k and k-1 suffice)
        if it > 0:
            st.X_prev, st.g_prev, st.Dfobj_prev, st.fobj_prev = \
                st.X_cur, st.g_cur, st.Dfobj_cur, st.fobj_cur
        st.X_cur, st.g_cur, st.Dfobj_cur, st.fobj_cur, st.f_cur = \
            X_it, g, Dfobj, fobj, f
        if o.scheme == "AMM":
            if it == 0:
                st.s[0] = 1.0
                st.oscillations.append(1)
            s0 = st.s[it]
            s1 = 0.5 + 0.5 * math.sqrt(4.0 * s0 * s0 + 1.0)
            st.s[it + 1] = s1
            st.gamma = (s0 - 1) / s1
            if fobj <= st.Fk[1]:
                st.soft_restart_hits[0] = st.soft_restart_hits[0] - 2 \
                    if st.soft_restart_hits[0] > 2 else 0
            else:
                st.soft_restart_hits[0] += 1
            if it > 0:
                if fobj <= st.fobj_prev:
                    st.soft_restart_hits[1] = 0
                    st.oscillations.append(1)
                else:
                    st.soft_restart_hits[1] += 1
                    st.oscillations.append(0)
                st.num_oscillations += int(st.oscillations[it] != st.oscillations[it - 1])
            if it > o.oscillation_cnt_period:
                k = it - o.oscillation_cnt_period
                st.num_oscillations -= int(st.oscillations[k] != st.oscillations[k - 1])
            st.Fk[0] = st.Fk[0] * (1 - o.eta[0]) + fobj * o.eta[0]
            st.Fk[1] = max(fobj, st.Fk[1] * (1 - o.eta[1]) + fobj * o.eta[1])
        else:
            st.Fk = [fobj, fobj]
        st.updated = True
        return 0

    # DPGOHash.cpp:230-444
    def amm_pgo(self):
        st, p, o = self.st, self.problem, self.opts
        assert st.updated
        n0, d = p.n[0], p.d
        it = st.iters
        if it == 0:
            Y = st.Xk.copy()
            g = st.g_cur.copy()
            Df = st.Dfobj_cur.copy()
        else:
            Y = st.X_cur + st.gamma * (st.X_cur - st.X_prev)
            if p.quadratic:
                g = st.g_cur + st.gamma * (st.g_cur - st.g_prev)
                Df = st.Dfobj_cur + st.gamma * (st.Dfobj_cur - st.Dfobj_prev)
            else:
                g, Df = p.evaluate_g_and_Df(Y)
        f = st.f_cur
        gk, fobj_k = st.g_cur, st.fobj_cur
        refined = (((st.gradFnorm * st.gradFnorm / fobj_k) > o.accepted_delta) or
                   (st.num_oscillations >= o.max_oscillations)) and \
            (o.max_iterations > 0) and (o.max_iterations_accepted > 0)
        st.last_refined = refined
        Fk = st.Fk
        st.Xakh = p.proximal(Y, Df)
        Gkh = p.evaluate_G(st.Xakh, gk, f)
        diff = st.Xakh - st.Xak
        minG = Fk[0] - o.psi * float(np.sum(diff * diff))
        st.Xak = st.Xak.copy()
        st.Xak[n0:] = st.Xakh[n0:]
        st.Xak[:n0] = p.recover_translations(st.Xak[n0:], g)
        if refined:
            res = self.ops.run_tnt(st.Xak, g, f, st)
            st.Xak = res.x
        st.Gk = p.evaluate_G(st.Xak, gk, f)
        if Gkh > minG:
            st.Xakh = p.proximal(st.Xk, st.Dfobj_cur)
            Gkh = p.evaluate_G(st.Xakh, gk, f)
        hard = st.Gk > Fk[0]
        soft = (st.Gk > Fk[1] and st.soft_restart_hits[0] >= o.max_soft_restart_hits[0]) or \
               (st.Gk > fobj_k and st.soft_restart_hits[1] > o.max_soft_restart_hits[1])
        if hard or soft:
            st.n_restarts += 1
            g = gk
            if Gkh <= fobj_k:
                st.Xak = st.Xakh.copy()
            else:
                st.Xak = p.proximal(st.Xk, st.Dfobj_cur)
            st.Xak[:n0] = p.recover_translations(st.Xak[n0:], gk)
            if refined:
                res = self.ops.run_tnt(st.Xak, g, f, st)
                st.Xak = res.x
                st.Gk = res.f
            else:
                st.Gk = p.evaluate_G(st.Xak, gk, f)
            if hard:
                st.s[it + 1] = max(0.5 * st.s[it + 1], 1.0)
            st.soft_restart_hits[0] //= 3
            st.soft_restart_hits[1] = 0
        if (Fk[0] - st.Gk) < o.phi * (Fk[0] - Gkh):
            st.Xak[n0:] = st.Xakh[n0:]
            st.Xak[:n0] = p.recover_translations(st.Xak[n0:], g)
            st.Gk = p.evaluate_G(st.Xak, gk, f)
        return 0

    # DPGOHash.cpp:446-581
    def mm_pgo(self):
        st, p, o = self.st, self.problem, self.opts
        n0 = p.n[0]
        g, Df, f = st.g_cur, st.Dfobj_cur, st.f_cur
        refined = ((st.gradFnorm * st.gradFnorm / st.fobj_cur) > o.accepted_delta) \
            and (o.max_iterations > 0) and (o.max_iterations_accepted > 0)
        st.last_refined = refined
        st.Xakh = p.proximal(st.Xk, Df)
        st.Xakh[:n0] = p.recover_translations(st.Xakh[n0:], g)
        if refined:
            res = self.ops.run_tnt(st.Xakh, g, f, st)
            st.Xak = res.x
            st.Gk = res.f
        else:
            st.Xak = st.Xakh.copy()
            st.Gk = p.evaluate_G(st.Xak, g, f)
        return 0

    # DPGOHash.cpp:583-628
    def iterate(self):
        st, p = self.st, self.problem
        if self.opts.scheme == "AMM":
            self.amm_pgo()
        else:
            self.mm_pgo()
        st.iters += 1
        st.Xk[:p.size0] = st.Xak
        st.updated = False
        return 0

    # DPGOHash.h:28-86
    def communicate(self, pgos):
        p, st = self.problem, self.st
        d, (n0, n1) = p.d, p.n
        base = (d + 1) * n0
        nbr_node, nbr_local = p.nbr_node, p.nbr_local
        ar = np.arange(d)
        for b in np.unique(nbr_node):
            q = pgos[int(b)]
            sel = np.nonzero(nbr_node == b)[0]
            j = nbr_local[sel]
            qn0 = q.problem.n[0]
            st.Xk[base + sel] = q.st.Xk[j]
            st.Xk[(base + n1 + d * sel[:, None] + ar).ravel()] = \
                q.st.Xk[(qn0 + d * j[:, None] + ar).ravel()]
        return 0


def build_comm_maps(problems, g_index):
    """Index arrays behind DPGO::communicate (DPGO_utils.h:397-453) and
    DPGOStar::communicate_n (DPGOStar.cpp:276-313): for every node the global
    ids of its own poses and of its neighbour copies, and where each
    neighbour lives in its owner's local numbering."""
    for a, p in enumerate(problems):
        p.own_gid = np.array([g_index[a][int(q)] for q in p.info.own_poses],
                             dtype=np.int64)
        keys = p.info.nbr_keys
        p.nbr_node = (keys >> 40).astype(np.int64)
        nbr_pose = (keys & ((np.int64(1) << 40) - 1)).astype(np.int64)
        p.nbr_local = np.empty(len(keys), dtype=np.int64)
        p.nbr_gid = np.empty(len(keys), dtype=np.int64)
        for b in np.unique(p.nbr_node):
            sel = p.nbr_node == b
            q = problems[int(b)]
            p.nbr_local[sel] = np.searchsorted(q.info.own_poses, nbr_pose[sel])
        p._nbr_pose = nbr_pose
    for a, p in enumerate(problems):
        for b in np.unique(p.nbr_node):
            sel = p.nbr_node == b
            p.nbr_gid[sel] = problems[int(b)].own_gid[p.nbr_local[sel]]


def gather_global(hashes, g_index, num_poses, d):
    """dist_pgo.cpp:502-511: global X = [t (N rows); R (dN rows)]."""
    X = np.zeros(((d + 1) * num_poses, d))
    for a, h in enumerate(hashes):
        n0 = h.problem.n[0]
        i = next(iter(g_index[a].values()))
        X[i:i + n0] = h.st.Xk[:n0]
        X[num_poses + d * i: num_poses + d * (i + n0)] = h.st.Xk[n0:(d + 1) * n0]
    return X


def scatter_initial(X, problems, g_index, num_poses, d):
    """dist_pgo.cpp:436-446 + DPGO::communicate (DPGO_utils.h:397-453): build
    each node's Z = [t;R;t_nbr;R_nbr] from a global X (needs build_comm_maps)."""
    out = []
    ar = np.arange(d)
    for a, p in enumerate(problems):
        n0, n1 = p.n
        Z = np.zeros(((d + 1) * (n0 + n1), d))
        og, ng = p.own_gid, p.nbr_gid
        Z[:n0] = X[og]
        Z[n0:(d + 1) * n0] = X[(num_poses + d * og[:, None] + ar).ravel()]
        base = (d + 1) * n0
        Z[base:base + n1] = X[ng]
        Z[base + n1:] = X[(num_poses + d * ng[:, None] + ar).ravel()]
        out.append(Z)
    return out


class GlobalObjective:
    """DPGOStar::evaluate_f / evaluate_grad (DPGOStar.cpp:713-829) on the
    global X = [t (N); R (dN)], evaluated edge by edge (each inter-node edge
    once, DPGO_utils.cpp:262-268)."""

    def __init__(self, num_poses, num_nodes, meas_global_ids, part, opts):
        self.N = num_poses
        self.d = part.d
        self.opts = opts
        self.meas = meas_global_ids     # Measurements with global pose ids
        self.inter = part.i_node != part.j_node

    def _residuals(self, X):
        N, d, m = self.N, self.d, self.meas
        t = X[:N]
        Y = X[N:].reshape(N, d, d)
        i, j = m.i_pose, m.j_pose
        rt = np.sqrt(m.tau)[:, None] * (t[i] - t[j] + np.einsum("ek,ekc->ec", m.t, Y[i]))
        rR = np.sqrt(m.kappa)[:, None, None] * (
            np.einsum("ecr,eck->erk", m.R, Y[i]) - Y[j])
        return rt, rR

    def evaluate_f(self, X):
        """DPGOStar.cpp:713-761.  Trivial loss: 1/2 tr(X^T M X) with the
        triplet form of M (DPGO_utils.cpp:500-560: kappa*I diagonal blocks,
        i.e. R_e^T R_e is NOT formed); robust: residual form via B0, B1."""
        if self.opts.loss == "trivial":
            return 0.5 * float(self._e_Mform(X).sum())
        rt, rR = self._residuals(X)
        e = np.square(rt).sum(axis=1) + np.square(rR).sum(axis=(1, 2))
        f = 0.5 * e[~self.inter].sum()
        _, fE = loss_weights(e[self.inter], self.opts.loss, self.opts.loss_reg)
        return float(f + fE)

    def _e_Mform(self, X):
        N, d, m = self.N, self.d, self.meas
        Y = X[N:].reshape(N, d, d)
        rt, _ = self._residuals(X)
        Yi, Yj = Y[m.i_pose], Y[m.j_pose]
        RtYi = np.einsum("ecr,eck->erk", m.R, Yi)
        rot = np.square(Yi).sum(axis=(1, 2)) + np.square(Yj).sum(axis=(1, 2)) \
            - 2.0 * (RtYi * Yj).sum(axis=(1, 2))
        return np.square(rt).sum(axis=1) + m.kappa * rot

    def evaluate_grad(self, X):
        """DPGOStar.cpp:763-829."""
        N, d, m = self.N, self.d, self.meas
        rt, rR = self._residuals(X)
        e = np.square(rt).sum(axis=1) + np.square(rR).sum(axis=(1, 2))
        w = np.ones(len(e))
        w[self.inter], _ = loss_weights(e[self.inter], self.opts.loss, self.opts.loss_reg)
        a = (w * np.sqrt(m.tau))[:, None] * rt
        B = (w * np.sqrt(m.kappa))[:, None, None] * rR
        Dt = np.zeros((N, d))
        DY = np.zeros((N, d, d))
        np.add.at(Dt, m.i_pose, a)
        np.add.at(Dt, m.j_pose, -a)
        if self.opts.loss == "trivial":
            Y = X[N:].reshape(N, d, d)
            RB = m.kappa[:, None, None] * (Y[m.i_pose] - np.einsum(
                "erc,eck->erk", m.R, Y[m.j_pose]))          # M-form
        else:
            RB = np.einsum("erc,eck->erk", m.R, B)
        np.add.at(DY, m.i_pose, m.t[:, :, None] * a[:, None, :] + RB)
        np.add.at(DY, m.j_pose, -B)
        Df = np.vstack([Dt, DY.reshape(N * d, d)])
        grad = Df.copy()
        grad[N:] = sod.proj(X[N:], Df[N:], d)
        return grad


class DPGOStar:
    """AMM-PGO* master-node driver (DPGOStar.cpp:126-711).  The global
    objective is delegated to GlobalObjective."""

    def __init__(self, num_nodes, per_node_meas, g_index, num_poses, gobj, opts):
        self.opts = opts
        self.num_nodes = num_nodes
        self.num_poses = num_poses
        self.g_index = g_index
        self.gobj = gobj
        self.problems = [DPGOProblem(a, per_node_meas[a], opts) for a in range(num_nodes)]
        self.ops = [_NodeOps(p, opts) for p in self.problems]
        build_comm_maps(self.problems, g_index)
        self.d = self.problems[0].d
        self.first = [next(iter(g.values())) for g in g_index]

    def _put(self, Xg, a, Xa):
        n0, d, N, i = self.problems[a].n[0], self.d, self.num_poses, self.first[a]
        Xg[i:i + n0] = Xa[:n0]
        Xg[N + d * i: N + d * (i + n0)] = Xa[n0:]

    # DPGOStar.cpp:109-124, :233-271
    def initialize(self, X):
        self.results = []
        Zs = scatter_initial(X, self.problems, self.g_index, self.num_poses, self.d)
        for a, p in enumerate(self.problems):
            st = NodeState()
            st.Xk = Zs[a]
            st.Xak = st.Xk[:p.size0].copy()
            st.Xakh = np.zeros_like(st.Xak)
            st.updated = False
            self.results.append(st)
        self.Xk = X.copy()
        self.Xkh = np.zeros_like(X)
        self.Xkp = np.zeros_like(X)
        self.fobj = self.gobj.evaluate_f(self.Xk)
        self.F = self.fobj
        return 0

    # DPGOStar.cpp:315-390
    def update(self):
        for a in range(self.num_nodes):
            self._update_n(a)
        return 0

    def _update_n(self, a):
        p = self.problems[a]
        st = self.results[a]
        if st.updated:
            return
        it = st.iters
        X_it = st.Xk.copy()
        if p.quadratic:
            g, f = p.evaluate_none_g_and_f0(X_it)
            fobj = p.evaluate_G(st.Xak, g, f)
            Dfobj = g + p.G @ st.Xak
        elif self.opts.rescale == "Static":
            g, f, Dfobj, fobj, st.DfobjE, st.fobjE = p.evaluate_g_and_f0(X_it)
        else:                                                   # DPGOStar.cpp:350-355
            g, f, Dfobj, fobj, st.DfobjE, st.fobjE, st.rescale_count = p.evaluate_g_and_f0_rescale(
                X_it, st.rescale_count, self.opts.max_rescale_count)
        st.Gk = fobj
        st.gradF = p.full_tangent_space_projection(st.Xak, Dfobj)
        st.gradFnorm = float(np.linalg.norm(st.gradF))
        if it > 0:
            st.X_prev, st.g_prev, st.Dfobj_prev = st.X_cur, st.g_cur, st.Dfobj_cur
        st.X_cur, st.g_cur, st.Dfobj_cur, st.fobj_cur, st.f_cur = X_it, g, Dfobj, fobj, f
        if self.opts.scheme == "AMM":
            if it == 0:
                st.s[0] = 1.0
            s0 = st.s[it]
            st.s[it + 1] = 0.5 + 0.5 * math.sqrt(4.0 * s0 * s0 + 1.0)
            st.gamma = (s0 - 1) / st.s[it + 1]
        st.Fk = [fobj, fobj]
        st.updated = True

    # DPGOStar.cpp:392-550
    def _amm_pgo_n(self, a):
        st, p, o = self.results[a], self.problems[a], self.opts
        n0 = p.n[0]
        it = st.iters
        if it == 0:
            Y, g, Df = st.Xk.copy(), st.g_cur.copy(), st.Dfobj_cur.copy()
        else:
            Y = st.X_cur + st.gamma * (st.X_cur - st.X_prev)
            if p.quadratic:
                g = st.g_cur + st.gamma * (st.g_cur - st.g_prev)
                Df = st.Dfobj_cur + st.gamma * (st.Dfobj_cur - st.Dfobj_prev)
            else:
                g, Df = p.evaluate_g_and_Df(Y)
        refined = (st.gradFnorm * st.gradFnorm / st.fobj_cur) > o.accepted_delta
        st.last_refined = refined
        st.Xakh = p.proximal(Y, Df)
        st.Xak = st.Xak.copy()
        st.Xak[n0:] = st.Xakh[n0:]
        st.Xak[:n0] = p.recover_translations(st.Xak[n0:], g)
        if refined:
            st.Xak = self.ops[a].run_tnt(st.Xak, g, st.f_cur, st).x
        self._put(self.Xkh, a, st.Xakh)
        self._put(self.Xkp, a, st.Xak)

    # DPGOStar.cpp:552-683
    def _mm_pgo_n(self, a):
        st, p, o = self.results[a], self.problems[a], self.opts
        n0 = p.n[0]
        g, f = st.g_cur, st.f_cur
        refined = (st.gradFnorm * st.gradFnorm / st.fobj_cur) > o.accepted_delta
        st.Xak = st.Xak.copy()
        st.Xak[n0:] = st.Xakh[n0:]
        st.Xak[:n0] = p.recover_translations(st.Xak[n0:], g)
        if refined:
            res = self.ops[a].run_tnt(st.Xak, g, f, st)
            st.Xak, st.Gk = res.x, res.f
        else:
            st.Gk = p.evaluate_G(st.Xak, g, f)
        self._put(self.Xkp, a, st.Xak)

    # DPGOStar.cpp:685-711
    def _pm_pgo_n(self, a):
        st, p = self.results[a], self.problems[a]
        st.Xakh = p.proximal(st.Xk, st.Dfobj_cur)
        self._put(self.Xkh, a, st.Xakh)

    def _restart_n(self, a):
        st = self.results[a]
        self._mm_pgo_n(a)
        st.s[st.iters + 1] = max(0.5 * st.s[st.iters + 1], 1.0)

    def _safeguard_n(self, a):
        st, p = self.results[a], self.problems[a]
        n0 = p.n[0]
        st.Xak[n0:] = st.Xakh[n0:]
        st.Xak[:n0] = p.recover_translations(st.Xak[n0:], st.g_cur)
        self._put(self.Xkp, a, st.Xak)

    def _finish_n(self, a):
        st, p = self.results[a], self.problems[a]
        st.iters += 1
        st.Xk[:p.size0] = st.Xak
        st.updated = False

    # DPGOStar.cpp:126-213
    def iterate(self):
        o = self.opts
        self.n_global_restarts = getattr(self, "n_global_restarts", 0)
        for a in range(self.num_nodes):
            self._amm_pgo_n(a)
        fobjh = self.gobj.evaluate_f(self.Xkh)
        if fobjh > self.F - o.psi * float(np.sum(np.square(self.Xkh - self.Xk))):
            for a in range(self.num_nodes):
                self._pm_pgo_n(a)
            fobjh = self.gobj.evaluate_f(self.Xkh)
        fobj = self.gobj.evaluate_f(self.Xkp)
        if fobj > self.F - o.psi * float(np.sum(np.square(self.Xkp - self.Xk))):
            self.n_global_restarts += 1
            for a in range(self.num_nodes):
                self._restart_n(a)
            fobj = self.gobj.evaluate_f(self.Xkp)
        if self.F - fobj < o.phi * (self.F - fobjh):
            for a in range(self.num_nodes):
                self._safeguard_n(a)
            fobj = self.gobj.evaluate_f(self.Xkp)
        for a in range(self.num_nodes):
            self._finish_n(a)
        self.Xk, self.Xkp = self.Xkp, self.Xk
        self.fobj = fobj
        self.F = self.F * (1 - o.eta[0]) + fobj * o.eta[0]
        return 0

    # DPGOStar.cpp:215-223, :276-313
    def communicate(self):
        Zs = scatter_initial(self.Xk, self.problems, self.g_index, self.num_poses, self.d)
        for a, p in enumerate(self.problems):
            st = self.results[a]
            st.Xk[p.size0:] = Zs[a][p.size0:]
            st.updated = False
        return 0

    def _communicate_n(self, a):
        """communicate() for one node (same values; used by the multi-process runner)."""
        p, st, d, N = self.problems[a], self.results[a], self.d, self.num_poses
        n0, n1 = p.n
        ng = p.nbr_gid
        base = (d + 1) * n0
        st.Xk[base:base + n1] = self.Xk[ng]
        st.Xk[base + n1:] = self.Xk[(N + d * ng[:, None] + np.arange(d)).ravel()]
        st.updated = False
