"""Generates the golden fixtures under tests/golden/ (run in the authoring container,
where /root/reference/dataset exists):

    python tests/golden/make_golden.py

For each case: the parsed measurement arrays of a reference dataset (derived data,
through dpgo_b200.read_g2o == DPGO::read_g2o_file), the initial iterate (centralised
chordal initialisation, what `dist_pgo --dist_init false` uses, dist_pgo.cpp:416-444),
and the ORACLE's per-iteration traces (per-node fobj, 2F, final poses, IRLS weights).
The `-m gpu` tests replay the CUDA path against these files on the GPU box, where
/root/reference does not exist; the `-m "not gpu"` tests check the oracle still
reproduces them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dpgo_b200 as D  # noqa: E402
from oracle import dist_pgo as odist  # noqa: E402
from oracle import dpgo as odpgo  # noqa: E402
from parity import to_measurements  # noqa: E402

DATA = "/root/reference/dataset"

CASES = [
    # name, file, nodes, loss, algorithm, iters, outlier fraction
    ("tinyGrid3D_n2_trivial_hash", "tinyGrid3D.g2o", 2, "trivial", "hash", 30, 0.0),
    ("smallGrid3D_n4_huber_star", "smallGrid3D.g2o", 4, "huber", "star", 30, 0.0),
    ("sphere2500_n4_trivial_hash", "sphere2500.g2o", 4, "trivial", "hash", 1000, 0.0),
    ("city10000_n16_gm_hash", "city10000.g2o", 16, "gm", "hash", 50, 0.1),
]


def inject_outliers(g, frac, seed=20241019):
    """BASELINE.json config 3: replace `frac` of the non-odometry edges by uniformly
    random SE(2) measurements (fixed seed)."""
    rng = np.random.default_rng(seed)
    lc = np.nonzero(np.abs(g.j.astype(np.int64) - g.i) != 1)[0]
    sel = np.sort(rng.permutation(len(lc))[: int(round(frac * len(lc)))])
    idx = lc[sel]
    th = rng.uniform(-np.pi, np.pi, len(idx))
    g.R[idx] = np.stack([np.stack([np.cos(th), -np.sin(th)], -1),
                         np.stack([np.sin(th), np.cos(th)], -1)], axis=1)
    g.t[idx] = rng.uniform(-10.0, 10.0, (len(idx), 2))
    return idx


def main():
    for name, fn, nn, loss, alg, iters, frac in CASES:
        g = D.read_g2o(os.path.join(DATA, fn))
        out_idx = inject_outliers(g, frac) if frac > 0 else np.zeros(0, dtype=np.int64)
        meas = to_measurements(g)
        X0 = odist.chordal_initialization(g.num_poses, meas)
        opts = odpgo.Options(loss=loss, preconditioner="BlockJacobi")
        res = odist.run(meas, g.num_poses, nn, opts, X0, iters, alg)
        # poses after the 50 iterations north_star names (final poses of longer runs are compared loosely:
        # weakly constrained directions let them drift at an objective agreement of 1e-8)
        X_50 = odist.run(meas, g.num_poses, nn, opts, X0, 50, alg, log_global=False)["X"] if iters > 50 else res["X"]
        w = [np.zeros(0) if x is None else x for x in res["weights"]]
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            d=g.d, num_poses=g.num_poses, num_nodes=nn, loss=loss, algorithm=alg, iters=iters,
            i=g.i, j=g.j, R=g.R, t=g.t, kappa=g.kappa, tau=g.tau, X0=X0,
            outlier_edges=out_idx,
            fobj_nodes=np.array(res["fobj_nodes"]), trace=np.array(res["trace"]),
            refined=np.array(res["refined"]), X_final=res["X"], X_50=X_50,
            weights=np.concatenate(w), weights_off=np.cumsum([0] + [len(x) for x in w]))
        print(name, "2F:", res["trace"][0][0], "->", res["trace"][-1][0])


if __name__ == "__main__":
    main()
