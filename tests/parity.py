"""Shared parity harness: run the oracle and the CUDA path on the same graph,
partition, initial iterate and iteration count, and compare traces.
(test infrastructure: the only place oracle and product meet)."""
from __future__ import annotations

import numpy as np

import dpgo_b200 as D
from oracle import dist_pgo as odist
from oracle import dpgo as odpgo
from oracle import g2o as og2o


def to_measurements(g):
    z = np.zeros(g.num_edges, dtype=np.int64)
    return og2o.Measurements(g.d, z, g.i.astype(np.int64), z.copy(), g.j.astype(np.int64),
                             g.R, g.t, g.kappa, g.tau)


LOSS_TO_ORACLE = {"trivial": "trivial", "huber": "huber", "gm": "gm", "welsch": "welsch"}


def run_both(g, num_nodes, X0, iters, loss="trivial", algorithm="hash", scheme="AMM",
             preconditioner="BlockJacobi", **kw):
    """Returns dict with oracle / cuda traces of 2F and per-node fobj."""
    oopts = odpgo.Options(loss=LOSS_TO_ORACLE[loss], scheme=scheme, preconditioner=preconditioner,
                          **{k: v for k, v in kw.items() if hasattr(odpgo.Options(), k)})
    meas = to_measurements(g)
    copts = D.Options(loss=loss, scheme=scheme, preconditioner=preconditioner,
                      **{k: v for k, v in kw.items()})
    cls = D.DPGOStar if algorithm == "star" else D.DPGOHash
    drv = cls(g, num_nodes, copts)
    if preconditioner == "RegularizedCholesky":
        # the regulariser is lambda_max(G11) / 1e6 with lambda_max a LOOSE estimate in the reference (Spectra, tolerance
        # 1e-4, DPGOProblem.cpp:114-118): both sides get the library's estimate, which is also checked against scipy's
        oopts.lambda_max_override = [drv.preconditioner_info(a)[0] for a in range(num_nodes)]
    ref = odist.run(meas, g.num_poses, num_nodes, oopts, X0, iters, algorithm)
    assert drv.initialize(X0) == 0
    assert drv.update() == 0
    fn = [[drv.node_scalars(a).fobj for a in range(num_nodes)]]
    refined = []
    for it in range(iters):
        rc = drv.iterate()
        assert rc == 0, D.load().mmpgo_last_error()
        refined.append([bool(drv.node_scalars(a).refined) for a in range(num_nodes)])
        assert drv.communicate() == 0
        rc = drv.update()
        assert rc == 0, D.load().mmpgo_last_error()
        fn.append([drv.node_scalars(a).fobj for a in range(num_nodes)])
    out = {
        "ref": ref, "drv": drv,
        "fobj_ref": np.array(ref["fobj_nodes"]), "fobj_gpu": np.array(fn),
        "refined_ref": np.array(ref["refined"]), "refined_gpu": np.array(refined),
        "X_ref": ref["X"], "X_gpu": drv.X(),
    }
    out["F_ref"] = np.array([t[0] for t in ref["trace"]]) / 2.0
    return out


def rel_trace_error(out):
    a, b = out["fobj_ref"].sum(axis=1), out["fobj_gpu"].sum(axis=1)
    return np.abs(a - b) / np.abs(a)


def pose_error(out, d):
    Xr, Xg = out["X_ref"], out["X_gpu"]
    N = Xr.shape[0] // (d + 1)
    et = np.abs(Xr[:N] - Xg[:N]).max()
    eR = np.sqrt(np.square((Xr[N:] - Xg[N:]).reshape(N, d * d)).sum(axis=1)).max()
    return et, eR
