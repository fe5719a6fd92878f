"""Driver loop of the `dist_pgo` example (oracle; test infrastructure only).

Restates C++/examples/dist_pgo.cpp:446-531 (initialize/update, then per
iteration: iterate all nodes -> gather global X -> communicate -> update ->
log 2F and 2||grad F||) and, for AMM-PGO*, the loop DPGOStar expects
(update; iterate; communicate, DPGOStar.cpp:126-231).  Also holds the
centralised chordal initialisation used when `--dist_init false`
(dist_pgo.cpp:416-444; SESync_utils.cpp:573-652), as host-side helper.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import dpgo, g2o, sod


def chordal_initialization(num_poses, meas):
    """Centralised chordal relaxation with pose 0 fixed to the identity,
    then least-squares translations.  Returns global X = [t (N); R^T blocks]."""
    N, d, m = num_poses, meas.d, len(meas)
    i, j = meas.i_pose, meas.j_pose
    # rotation: minimise sum kappa || R_e^T Y_i - Y_j ||^2, Y_0 = I
    rows, cols, vals = [], [], []
    e = np.arange(m)
    sk = np.sqrt(meas.kappa)
    for r in range(d):
        for c in range(d):
            rows.append(d * e + r); cols.append(d * i + c); vals.append(sk * meas.R[:, c, r])
        rows.append(d * e + r); cols.append(d * j + r); vals.append(-sk)
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(d * m, d * N)).tocsc()
    A0, A1 = A[:, :d], A[:, d:]
    rhs = -(A0 @ np.eye(d))
    Yr = spla.spsolve((A1.T @ A1).tocsc(), A1.T @ rhs)
    Y = np.vstack([np.eye(d), np.asarray(Yr).reshape(-1, d)])
    Y = sod.project_svd(Y.reshape(N, d, d)).reshape(N * d, d)
    # translations: minimise sum tau || t_i - t_j + t_e^T Y_i ||^2, t_0 = 0
    st = np.sqrt(meas.tau)
    Bt = sp.coo_matrix((np.concatenate([st, -st]), (np.concatenate([e, e]),
                                                     np.concatenate([i, j]))),
                       shape=(m, N)).tocsc()
    c = st[:, None] * np.einsum("ek,ekc->ec", meas.t, Y.reshape(N, d, d)[i])
    B1 = Bt[:, 1:]
    tr = spla.spsolve((B1.T @ B1).tocsc(), -(B1.T @ c))
    t = np.vstack([np.zeros((1, d)), np.asarray(tr).reshape(-1, d)])
    return np.vstack([t, Y])


def odometry_initialization(num_poses, meas):
    """Chain the i -> i+1 measurements from pose 0 (identity)."""
    N, d = num_poses, meas.d
    t = np.zeros((N, d))
    Y = np.tile(np.eye(d), (N, 1, 1))          # Y_i = R_i^T
    odo = {int(a): k for k, (a, b) in enumerate(zip(meas.i_pose, meas.j_pose)) if b == a + 1}
    for a in range(N - 1):
        k = odo.get(a)
        if k is None:
            t[a + 1], Y[a + 1] = t[a], Y[a]
            continue
        Ri = Y[a].T
        t[a + 1] = t[a] + Ri @ meas.t[k]
        Y[a + 1] = (Ri @ meas.R[k]).T
    return np.vstack([t, Y.reshape(N * d, d)])


def run(meas, num_poses, num_nodes, opts, X0, iters, algorithm="hash",
        log_global=True, timing=None, workers=1):
    """Returns dict(trace=[(2F, 2|grad|)...], X=global X, per-node scalars).
    `timing`, if a dict, receives seconds spent in iterate/update/communicate
    (what dist_pgo.cpp:496-521 times, plus communicate).  `workers` > 1 (AMM-PGO* only) spreads the
    per-node work over forked processes (oracle/parallel.py; same iterates, CPU-baseline timing)."""
    per_node, g_index, part = g2o.partition(num_poses, num_nodes, meas)
    d = meas.d
    gobj = dpgo.GlobalObjective(num_poses, num_nodes, meas, part, opts)
    out = {"trace": [], "fobj_nodes": [], "refined": [], "tcg": []}
    t_acc = 0.0
    if algorithm == "hash":
        hashes = [dpgo.DPGOHash(a, per_node[a], opts) for a in range(num_nodes)]
        problems = [h.problem for h in hashes]
        dpgo.build_comm_maps(problems, g_index)
        Zs = dpgo.scatter_initial(X0, problems, g_index, num_poses, d)
        for h, Z in zip(hashes, Zs):
            h.initialize(Z)
            h.update()
        X = dpgo.gather_global(hashes, g_index, num_poses, d)

        def log():
            if log_global:
                F = gobj.evaluate_f(X)
                gn = float(np.linalg.norm(gobj.evaluate_grad(X)))
                out["trace"].append((2 * F, 2 * gn))
            out["fobj_nodes"].append([h.st.fobj_cur for h in hashes])
        log()
        for it in range(iters):
            t0 = time.perf_counter()
            for h in hashes:
                h.iterate()
            out["refined"].append([h.st.last_refined for h in hashes])
            X = dpgo.gather_global(hashes, g_index, num_poses, d)
            for h in hashes:
                h.communicate(hashes)
            for h in hashes:
                h.update()
            t_acc += time.perf_counter() - t0
            log()
        out["tcg"] = [h.st.tcg_iters for h in hashes]
        out["restarts"] = [h.st.n_restarts for h in hashes]
        out["weights"] = [getattr(h.problem, "last_weights", None) for h in hashes]
        out["hashes"] = hashes
    elif algorithm == "star":
        if workers > 1:
            from . import parallel
            star = parallel.ParallelDPGOStar(num_nodes, per_node, g_index, num_poses, gobj, opts, workers=workers)
        else:
            star = dpgo.DPGOStar(num_nodes, per_node, g_index, num_poses, gobj, opts)
        star.initialize(X0)
        X = X0
        for it in range(iters + 1):
            t0 = time.perf_counter()
            star.update()
            t_acc += time.perf_counter() - t0
            if log_global:
                gn = float(np.linalg.norm(gobj.evaluate_grad(star.Xk)))
                out["trace"].append((2 * gobj.evaluate_f(star.Xk), 2 * gn))
            out["fobj_nodes"].append([st.fobj_cur for st in star.results])
            if it == iters:
                break
            t0 = time.perf_counter()
            star.iterate()
            star.communicate()
            t_acc += time.perf_counter() - t0
            out["refined"].append([st.last_refined for st in star.results])
        X = np.array(star.Xk)
        if workers > 1:
            star.close()
        out["tcg"] = [st.tcg_iters for st in star.results]
        out["restarts"] = star.n_global_restarts if hasattr(star, "n_global_restarts") else 0
        out["weights"] = [getattr(p, "last_weights", None) for p in star.problems]
        out["star"] = star
    else:
        raise ValueError(algorithm)
    out["X"] = X
    out["seconds"] = t_acc
    out["g_index"] = g_index
    if timing is not None:
        timing["seconds"] = t_acc
    return out
