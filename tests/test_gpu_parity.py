"""Parity of the CUDA path (through the C ABI) against the oracle and the golden
fixtures.  Tolerances: per-iteration objective 1e-8 relative, final poses 1e-6
(BASELINE.json north_star); outlier classification identical."""
import numpy as np
import pytest

import dpgo_b200 as D
import parity
from golden_util import golden_cases, load_golden

pytestmark = pytest.mark.gpu

F_TOL = 1e-8
POSE_TOL = 1e-6


def _check(out, d, iters_checked=None):
    err = parity.rel_trace_error(out)
    if iters_checked:
        err = err[: iters_checked + 1]
    assert err.max() < F_TOL, err
    et, eR = parity.pose_error(out, d)
    assert et < POSE_TOL and eR < POSE_TOL
    assert (out["refined_ref"] == out["refined_gpu"]).all()


@pytest.fixture(scope="module")
def grid():
    return D.grid3d(6, 6, 6, seed=1)


@pytest.mark.parametrize("loss", ["trivial", "huber", "gm", "welsch"])
@pytest.mark.parametrize("alg", ["hash", "star"])
def test_amm_parity_se3(grid, loss, alg):
    g, _, X0 = grid
    _check(parity.run_both(g, 4, X0, 12, loss=loss, algorithm=alg), 3)


@pytest.mark.parametrize("pre", ["None", "Jacobi", "BlockJacobi"])
def test_preconditioners(grid, pre):
    g, _, X0 = grid
    _check(parity.run_both(g, 4, X0, 8, preconditioner=pre), 3)


def test_mm_scheme(grid):
    g, _, X0 = grid
    _check(parity.run_both(g, 4, X0, 8, scheme="MM"), 3)


def test_no_refinement(grid):
    g, _, X0 = grid
    _check(parity.run_both(g, 4, X0, 8, max_iterations=0), 3)


@pytest.mark.parametrize("solver", ["direct", "pcg"])
def test_sparse_translation_solve(grid, solver):
    # nodes too large for a dense G00^{-1}: sparse Cholesky sweeps (default) or the Jacobi-PCG solve
    g, _, X0 = grid
    out = parity.run_both(g, 4, X0, 8, dense_solve_max_n=0, translation_solver=solver)
    _check(out, 3)
    assert out["drv"].solver_info()["solver"] == solver


@pytest.mark.parametrize("loss", ["trivial", "gm"])
def test_se2_with_outliers(loss):
    g, _, X0 = D.city2d(14, 12, seed=2)
    out = parity.run_both(g, 4, X0, 12, loss=loss)
    _check(out, 2)


def test_config2_grid3d_standin():
    """BASELINE.json configs[1]: grid3D (8000 poses, dense loop closures), 8 nodes, Huber, AMM-PGO*.
    dataset/grid3D.g2o is not shipped with the reference; a seeded 20 x 20 x 20 grid in the style of
    smallGrid3D.g2o stands in."""
    g, _, X0 = D.grid3d(20, 20, 20, seed=2)
    _check(parity.run_both(g, 8, X0, 10, loss="huber", algorithm="star"), 3)


def test_ragged_partition():
    # N not divisible by the node count (DPGO_utils.cpp:147-158)
    g, _, X0 = D.grid3d(5, 5, 3, seed=5)        # 75 poses over 4 nodes: 19,19,19,18
    _check(parity.run_both(g, 4, X0, 6), 3)
    assert [D.DPGOHash(g, 4).node_scalars(a).n0 for a in range(4)] == [19, 19, 19, 18]


def test_single_node_graph():
    # degenerate: one node, no inter-node edges.  G00 is then a pure graph Laplacian + xi I
    # (condition number ~1e13), so the translation solve amplifies rounding in the reference as
    # much as here; only a loose agreement is meaningful.
    g, _, X0 = D.grid3d(5, 5, 3, seed=5)
    out = parity.run_both(g, 1, X0, 4)
    err = parity.rel_trace_error(out)
    assert err.max() < 1e-2
    f = out["fobj_gpu"].sum(axis=1)
    assert f[-1] < 0.5 * f[0]


def test_sphere_rings_welsch():
    g, _, X0 = D.sphere_rings(6, 40, seed=3)
    _check(parity.run_both(g, 6, X0, 10, loss="welsch"), 3)


def test_fifty_iterations_trace(grid):
    # north_star: objective traces agree to 1e-8 relative for the first 50 iterations
    g, _, X0 = grid
    _check(parity.run_both(g, 4, X0, 50), 3, iters_checked=50)


@pytest.mark.parametrize("loss", ["huber", "gm", "welsch"])
def test_outlier_classification_identical(loss):
    g, _, X0 = D.city2d(14, 12, outlier_fraction=0.2, seed=9)
    out = parity.run_both(g, 4, X0, 10, loss=loss)
    thr = 1.0 if loss == "huber" else 0.5       # Huber: e > delta <=> w < 1; GM/Welsch: w < 1/2
    for a in range(4):
        w_ref = out["ref"]["weights"][a]
        w_gpu = out["drv"].weights(a)
        assert len(w_ref) == len(w_gpu)
        assert np.array_equal(w_ref < thr, w_gpu < thr)
        assert np.abs(w_ref - w_gpu).max() < 1e-9


@pytest.mark.parametrize("name", golden_cases())
def test_golden_fixture(name):
    g, z = load_golden(name)
    alg, loss, nn, iters = str(z["algorithm"]), str(z["loss"]), int(z["num_nodes"]), int(z["iters"])
    opts = D.Options(loss=loss, preconditioner="BlockJacobi")
    cls = D.DPGOStar if alg == "star" else D.DPGOHash
    drv = cls(g, nn, opts)
    assert drv.initialize(z["X0"]) == 0 and drv.update() == 0
    f = [sum(drv.node_scalars(a).fobj for a in range(nn))]
    X50 = None
    for it in range(iters):
        assert drv.iterate() == 0 and drv.communicate() == 0 and drv.update() == 0
        f.append(sum(drv.node_scalars(a).fobj for a in range(nn)))
        if it + 1 == 50 and iters > 50:
            X50 = drv.X()
    want = z["fobj_nodes"].sum(axis=1)
    err = np.abs(np.array(f) - want) / np.abs(want)
    # north_star: 1e-8 relative over the first 50 iterations; rounding differences then accumulate slowly
    # (sphere2500 runs its full 1000 iterations: 1.2e-8 at the end)
    assert err[:51].max() <= F_TOL and err.max() <= 1e-6, (err[:51].max(), err.max())
    X = drv.X()
    if X50 is None:
        assert np.abs(X - z["X_final"]).max() < POSE_TOL
    else:
        # 1e-6 after the 50 iterations north_star names; after 1000 the poses have drifted along the weakly
        # constrained directions of the graph (1.4e-4 here) while the objective still agrees to 1.2e-8
        assert np.abs(X50 - z["X_50"]).max() < POSE_TOL
        assert np.abs(X - z["X_final"]).max() < 1e-2
    if loss != "trivial":
        off = z["weights_off"]
        thr = 1.0 if loss == "huber" else 0.5
        for a in range(nn):
            w_ref = z["weights"][off[a]:off[a + 1]]
            w_gpu = drv.weights(a)
            assert np.array_equal(w_ref < thr, w_gpu < thr)


def test_evaluate_f_matches_oracle(grid):
    from oracle import dpgo as odpgo, g2o as og2o
    g, _, X0 = grid
    meas = parity.to_measurements(g)
    for loss in ("trivial", "huber", "gm", "welsch"):
        _, _, part = og2o.partition(g.num_poses, 4, meas)
        gobj = odpgo.GlobalObjective(g.num_poses, 4, meas, part, odpgo.Options(loss=loss))
        drv = D.DPGOStar(g, 4, D.Options(loss=loss))
        f = drv.evaluate_f(X0)
        want = gobj.evaluate_f(X0)
        assert abs(f - want) <= 1e-12 * abs(want)


@pytest.mark.parametrize("case", ["se3", "se2"])
def test_evaluate_grad_matches_oracle(grid, case):
    """DPGOStar::evaluate_grad (DPGOStar.cpp:763-829) at a perturbed iterate, all four losses; the
    call must not disturb a solve in progress."""
    from oracle import dpgo as odpgo, g2o as og2o
    g, _, X0 = grid if case == "se3" else D.city2d(14, 12, outlier_fraction=0.2, seed=3)
    meas = parity.to_measurements(g)
    for loss in ("trivial", "huber", "gm", "welsch"):
        _, _, part = og2o.partition(g.num_poses, 4, meas)
        gobj = odpgo.GlobalObjective(g.num_poses, 4, meas, part, odpgo.Options(loss=loss))
        drv = D.DPGOStar(g, 4, D.Options(loss=loss))
        assert drv.initialize(X0) == 0 and drv.update() == 0 and drv.iterate() == 0
        assert drv.communicate() == 0 and drv.update() == 0
        f_before = drv.objective()
        X1 = drv.X()
        got, want = drv.evaluate_grad(X1), gobj.evaluate_grad(X1)
        assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
        # |grad F| from the per-node sums of update() is the norm of the same gradient
        assert abs(np.linalg.norm(got) - f_before[1]) <= 1e-9 * max(1.0, f_before[1])
        assert drv.objective() == f_before and drv.iterate() == 0


def test_error_codes(grid):
    g, _, X0 = grid
    drv = D.DPGOHash(g, 4)
    assert drv.iterate() == -3                  # MMPGO_ERR_STATE: not initialized
    assert drv.initialize(X0[:-1]) == -1        # inconsistent size -> -1 like the reference
    assert drv.initialize(X0) == 0
    assert drv.iterate() == -3                  # iterate before update
    assert drv.update() == 0 and drv.iterate() == 0


def test_full_size_properties():
    """Size-independent properties on a larger graph than the oracle is run on:
    (i) sum of per-node objectives equals the edge-parallel global objective,
    (ii) MM-PGO is monotone, (iii) rotations stay on SO(3)."""
    g, _, X0 = D.grid3d(40, 40, 25, seed=11)     # 40k poses, ~160k edges, PCG solve path
    drv = D.DPGOHash(g, 8, D.Options(scheme="MM", dense_solve_max_n=0, translation_solver="pcg"))
    assert drv.initialize(X0) == 0 and drv.update() == 0
    prev = None
    for _ in range(5):
        f, _g = drv.objective()
        assert abs(f - drv.evaluate_f(drv.X())) <= 1e-9 * abs(f)
        if prev is not None:
            assert f <= prev * (1 + 1e-12)
        prev = f
        assert drv.iterate() == 0 and drv.communicate() == 0 and drv.update() == 0
    X = drv.X()
    N = g.num_poses
    Y = X[N:].reshape(N, 3, 3)
    assert np.abs(Y @ np.swapaxes(Y, 1, 2) - np.eye(3)).max() < 1e-12
    assert np.abs(np.linalg.det(Y) - 1).max() < 1e-12


@pytest.mark.parametrize("d", [2, 3])
def test_projection_bitwise_vs_reference_avx2(d):
    """The device polar projection against the outputs of the REAL reference kernels
    (project_to_SOd.cpp compiled from /root/reference with -ffp-contract=off; fixture made by
    tests/golden/make_projection_golden.py): bit-identical, since the device code spells the same
    mul / add / fma sequence.  The stock (FMA-contracted) reference build differs by ulps."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref", "sod_projection.npz"))
    A, want = z["A%d" % d], z["U%d" % d]
    got = D.project_to_SOdn(A)
    assert np.array_equal(got, want), np.abs(got - want).max()
    assert np.abs(got - z["U%d_fma" % d]).max() < 2e-14
    from oracle import ref
    if ref.available():                       # the compiled reference travels with the repo
        assert np.array_equal(ref.project(A, nofma=True), want)


# sparse Cholesky sweeps (default) / PCG copy-ring kernel (large shards) / PCG rendezvous-lean kernel (small shards)
TS_KERNELS = ["direct", "pcg_ring", "pcg_lite"]


@pytest.mark.parametrize("kernel", TS_KERNELS)
@pytest.mark.parametrize("alg,loss", [("hash", "trivial"), ("star", "huber")])
def test_persistent_solve_multi_tile_nodes(alg, loss, kernel):
    """Nodes spanning many CTA tiles and several chunks of the persistent translation solve
    (1600 poses per node, ragged last tile), masked sub-sets of nodes in the restart paths."""
    g, _, X0 = D.grid3d(20, 20, 12, seed=8)
    _check(parity.run_both(g, 3, X0, 6, loss=loss, algorithm=alg, dense_solve_max_n=0, translation_solver=kernel), 3)


@pytest.mark.parametrize("kernel", TS_KERNELS)
def test_persistent_solve_se2(kernel):
    g, _, X0 = D.city2d(40, 30, seed=6)
    _check(parity.run_both(g, 5, X0, 6, loss="gm", dense_solve_max_n=0, translation_solver=kernel), 2)


@pytest.mark.parametrize("kernel", TS_KERNELS)
def test_persistent_solve_many_small_nodes(kernel):
    # more nodes than a CTA has segments to spare: 40 nodes of 45 poses, one tile each
    g, _, X0 = D.grid3d(15, 12, 10, seed=12)
    _check(parity.run_both(g, 40, X0, 5, dense_solve_max_n=0, translation_solver=kernel), 3)


def _run_star(g, X0, nodes, iters, solver="pcg"):
    drv = D.DPGOStar(g, nodes, D.Options(loss="trivial", dense_solve_max_n=0, translation_solver=solver))
    assert drv.initialize(X0) == 0 and drv.update() == 0
    for _ in range(iters):
        assert drv.iterate() == 0, D.load().mmpgo_last_error()
        assert drv.communicate() == 0 and drv.update() == 0
    c = drv.counters()
    return drv.X(), drv.objective()[0], c.solve_iters


@pytest.mark.parametrize("dims,nodes", [((30, 30, 30), 5), ((24, 24, 16), 3)])
def test_tsolve_lite_bitwise_equals_ring(dims, nodes):
    """The two translation-solve kernels share the data layout, the per-pose arithmetic and the
    fixed-order reductions: same iterates to the last bit, same iteration counts.  The first case
    gives every lite CTA two CTA tiles, some straddling a node boundary (two segments per CTA)."""
    g, _, X0 = D.grid3d(*dims, seed=21)
    out = {}
    for kernel in ("ring", "lite"):
        out[kernel] = _run_star(g, X0, nodes, 3, "pcg_" + kernel)
    assert out["ring"][2] == out["lite"][2] and out["ring"][2] > 0
    assert out["ring"][1] == out["lite"][1]
    assert np.array_equal(out["ring"][0], out["lite"][0]), np.abs(out["ring"][0] - out["lite"][0]).max()


@pytest.mark.parametrize("alg,loss", [("hash", "gm"), ("star", "welsch")])
def test_fifty_iterations_trace_robust(alg, loss):
    # 50 iterations with a robust loss and 20 % outlier loop closures (BASELINE.json config 3 in small)
    g, _, X0 = D.city2d(14, 12, outlier_fraction=0.2, seed=9)
    _check(parity.run_both(g, 4, X0, 50, loss=loss, algorithm=alg), 2, iters_checked=50)


def test_poses_absent_from_every_measurement():
    """generate_data_info indexes only poses that occur in a measurement (DPGO_utils.cpp:358-372):
    ids without edges stay untouched in X, and the partition still counts them."""
    g, _, X0 = D.grid3d(5, 5, 4, seed=21)
    # renumber: insert two unused ids (7 and 60) by shifting the poses above them
    def shift(v):
        v = v.astype(np.int64)
        v = v + (v >= 7)
        return v + (v >= 60)
    N2 = g.num_poses + 2
    g2 = D.PoseGraph(3, N2, shift(g.i), shift(g.j), g.R, g.t, g.kappa, g.tau)
    keep = np.array([k for k in range(N2) if k not in (7, 60)])
    X2 = np.zeros((4 * N2, 3))
    X2[keep] = X0[:g.num_poses]
    X2[7], X2[60] = 123.0, -5.0                                    # sentinels in the unused rows
    for r in range(3):
        X2[N2 + 3 * keep + r] = X0[g.num_poses + 3 * np.arange(g.num_poses) + r]
        X2[N2 + 3 * np.array([7, 60]) + r, r] = 1.0
    out = parity.run_both(g2, 4, X2, 5)
    X = out["X_gpu"].copy()
    # the driver leaves rows of unused ids as the caller passed them; dist_pgo's own assembly of X
    # starts from zeros (dist_pgo.cpp:475), which is what the oracle returns there
    assert (X[7] == 123.0).all() and (X[60] == -5.0).all()
    # (the reference addresses a node's rows as [first id, first id + n0), DPGOStar.cpp:541-547,
    # which misplaces rows when ids are missing inside a node; only the objective is comparable)
    assert parity.rel_trace_error(out).max() < F_TOL
    assert (out["refined_ref"] == out["refined_gpu"]).all()
    Y = X[N2:].reshape(N2, 3, 3)[keep]
    assert np.abs(Y @ np.swapaxes(Y, 1, 2) - np.eye(3)).max() < 1e-12


def test_set_graph_rejects_bad_input():
    g, _, _ = D.grid3d(4, 4, 2, seed=3)
    bad = D.PoseGraph(3, g.num_poses, g.i.copy(), g.j.copy(), g.R, g.t, g.kappa, g.tau)
    bad.j[5] = g.num_poses + 3                    # endpoint out of range
    with pytest.raises(D.MmpgoError):
        D.DPGOHash(bad, 2)
    with pytest.raises(D.MmpgoError):
        D.DPGOHash(g, g.num_poses + 1)            # more nodes than poses
    # a node without any measurement (the reference would build an empty problem and fail later)
    iso = D.PoseGraph(3, 64, g.i % 16, (g.j % 16 + 1) % 16, g.R, g.t, g.kappa, g.tau)
    keep = iso.i != iso.j
    iso = D.PoseGraph(3, 64, iso.i[keep], iso.j[keep], g.R[keep], g.t[keep], g.kappa[keep], g.tau[keep])
    with pytest.raises(D.MmpgoError):
        D.DPGOHash(iso, 4)


# ---------------------------------------------------------------------------------------------
# round 2: parity at the benchmarked node size, several handles per run, driver call order
# ---------------------------------------------------------------------------------------------
def test_config3_node_size_slab():
    """BASELINE.json configs[3] per-node size: a 100 x 125 x 5 slab of the benchmark's grid generator =
    4 robot nodes of 15 625 poses each (the 1 M-pose workload has 64 of them), AMM-PGO*, trivial loss,
    sparse translation solve (no dense G00 inverse at this size), against the oracle."""
    g, _, X0 = D.grid3d(100, 125, 5)
    out = parity.run_both(g, 4, X0, 4, loss="trivial", algorithm="star")
    _check(out, 3)
    assert out["drv"].solver_info()["solver"] == "direct"
    assert [out["drv"].node_scalars(a).n0 for a in range(4)] == [15625] * 4


def test_config2_fifty_iterations():
    """configs[1] stand-in (see test_config2_grid3d_standin) over the 50 iterations north_star names."""
    g, _, X0 = D.grid3d(20, 20, 20, seed=2)
    _check(parity.run_both(g, 8, X0, 50, loss="huber", algorithm="star"), 3, iters_checked=50)


@pytest.mark.parametrize("alg,loss,dense", [("hash", "trivial", 2048), ("hash", "huber", 0), ("star", "trivial", 0),
                                            ("star", "welsch", 2048)])
@pytest.mark.parametrize("W", [2, 3])
def test_sharded_handles_reproduce_single_handle(alg, loss, dense, W):
    """W handles (one per rank of a W-GPU run) driven in lock step through the transport callbacks of
    mmpgo_set_sharding reproduce the single-handle run: same exchange plan, packing kernels and halo
    rows as the NCCL run, served in-process (tests/inproc_world.py) so that it runs on one GPU.
    AMM-PGO# has no scalar collective: bit-identical.  AMM-PGO* sums its four global scalars in a
    different association (per-rank partial sums), hence the tolerance."""
    from inproc_world import InProcWorld
    g, _, X0 = D.grid3d(12, 12, 12, seed=4)
    nodes, iters = 6, 10
    ref, tr = D.run_dist_pgo(g, nodes, X0, iters, D.Options(loss=loss, dense_solve_max_n=dense), alg)
    tr = np.array([t[0] / 2 for t in tr])
    world = InProcWorld(g, nodes, W, alg, loss=loss, dense_solve_max_n=dense)
    trace, X = world.run(X0, iters)
    assert all(rk.exchanges > 0 for rk in world.ranks)
    if alg == "hash":
        assert np.array_equal(X, ref.X()), np.abs(X - ref.X()).max()
        assert np.abs(trace - tr).max() <= 1e-13 * np.abs(tr).max()
    else:
        assert all(rk.allreduces > 0 for rk in world.ranks)
        assert np.abs(trace - tr).max() <= 1e-9 * np.abs(tr).max()
        assert np.abs(X - ref.X()).max() < 1e-7


def test_communicate_may_be_repeated_or_skipped(grid):
    """The reference's iterate() publishes X^{k+1} itself (DPGOHash.cpp:612-616); communicate() only
    refreshes neighbour copies and may be called twice, or not at all when every neighbour lives in
    the same handle."""
    g, _, X0 = grid
    def run(pattern):
        drv = D.DPGOHash(g, 4, D.Options(loss="huber"))
        assert drv.initialize(X0) == 0 and drv.update() == 0
        assert drv.communicate() == 0                         # before the first iterate: harmless
        for _ in range(5):
            assert drv.iterate() == 0
            for _ in range(pattern):
                assert drv.communicate() == 0
            assert drv.update() == 0
        return drv.X(), drv.objective()[0]
    X1, f1 = run(1)
    for pattern in (0, 2):
        X, f = run(pattern)
        assert f == f1 and np.array_equal(X, X1)
    # results().Xk right after iterate() is the new iterate
    drv = D.DPGOHash(g, 4)
    assert drv.initialize(X0) == 0 and drv.update() == 0 and drv.iterate() == 0
    assert np.abs(drv.X() - X0).max() > 1e-3
    assert drv.iterate() == -3                                # MMPGO_ERR_STATE: update() first


def _g00(g, nodes, xi=1e-11):
    """G00 of every node (DPGO_utils.cpp:2212-2243), block diagonal over nodes, scipy CSR."""
    import scipy.sparse as sp
    N = g.num_poses
    node = np.minimum(np.arange(N) // (N // nodes), nodes - 1) if N % nodes == 0 else None
    assert node is not None
    ii, jj = g.i.astype(np.int64), g.j.astype(np.int64)
    intra = node[ii] == node[jj]
    A = sp.coo_matrix((np.concatenate([-g.tau[intra]] * 2), (np.concatenate([ii[intra], jj[intra]]),
                                                             np.concatenate([jj[intra], ii[intra]]))), shape=(N, N)).tocsr()
    diag = np.full(N, xi)
    np.add.at(diag, ii[intra], g.tau[intra])
    np.add.at(diag, jj[intra], g.tau[intra])
    np.add.at(diag, ii[~intra], 2 * g.tau[~intra])
    np.add.at(diag, jj[~intra], 2 * g.tau[~intra])
    return (A + sp.diags(diag)).tocsr()


@pytest.mark.parametrize("solver,tol", [("direct", 1e-12), ("pcg", 1e-10)])
@pytest.mark.parametrize("case", ["grid", "slab", "rings", "se2"])
def test_translation_solve_against_scipy(solver, tol, case):
    """The linear solve of recover_translations (L_.solve, DPGOProblem.h:291) on random right-hand sides,
    every solve path against scipy's sparse LU: grid nodes of a few CTA tiles, the 15 625-pose slabs of the
    benchmark (fronts served by whole CTAs), ring-shaped robots (separators of two poses), SE(2)."""
    import scipy.sparse.linalg as spla
    if case == "grid":
        (g, _, _), nodes = D.grid3d(20, 20, 12, seed=8), 3
    elif case == "slab":
        (g, _, _), nodes = D.grid3d(100, 125, 5), 4
    elif case == "rings":
        (g, _, _), nodes = D.sphere_rings(6, 3000, seed=3), 6
    else:
        (g, _, _), nodes = D.city2d(60, 50, seed=6), 5
    drv = D.DPGOHash(g, nodes, D.Options(dense_solve_max_n=0, translation_solver=solver))
    assert drv.solver_info()["solver"] == solver
    A = _g00(g, nodes)
    rng = np.random.default_rng(5)
    b = rng.standard_normal((g.num_poses, g.d)) * 100.0
    t = drv.translation_solve(b)
    want = -spla.splu(A.tocsc()).solve(b)
    assert np.abs(t - want).max() <= tol * np.abs(want).max()
    # deterministic: a second solve returns the same bits
    assert np.array_equal(t, drv.translation_solve(b))


def test_direct_solve_bitwise_equals_host_restatement():
    """The device sweeps accumulate in the order of mf_host_solve (mmpgo_factor.cu) with fused multiply-adds:
    same bits as the host restatement of the same factor."""
    import ctypes as C
    import scipy.sparse as sp
    from dpgo_b200 import lib as L
    g, _, _ = D.grid3d(24, 24, 16, seed=21)
    nodes = 1
    drv = D.DPGOHash(g, nodes, D.Options(dense_solve_max_n=0, translation_solver="direct", regularizer=1e-3))
    A = sp.csr_matrix(_g00(g, nodes, xi=1e-3))
    A.sort_indices()
    # the ordering depends on the graph only and every (row, column) pair occurs once in this grid, so the
    # library's factor of G00 and the one built here from scipy's CSR hold the same numbers
    rng = np.random.default_rng(6)
    b = rng.standard_normal((g.num_poses, 3))
    x = np.zeros_like(b)
    ptr, col = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    L.check(L.load().mmpgo_mf_host_solve(A.shape[0], L.iptr(ptr), L.iptr(col), L.dptr(np.ascontiguousarray(A.data)), 1, 32, 3,
                                         L.dptr(b), L.dptr(x), None))
    t = drv.translation_solve(b)
    assert np.array_equal(t, -x), np.abs(t + x).max()


# ---------------------------------------------------------------------------------------------
# Rescale::Dynamic (DPGOProblem.cpp:289-358, 426-514, 751-840)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("loss", ["huber", "gm", "welsch"])
@pytest.mark.parametrize("alg", ["hash", "star"])
def test_dynamic_rescale_parity(alg, loss):
    """The per-measurement rescale vector clamp(1.25 omega, 0.01, 1) replaces the unit weights of the majoriser's
    inter-node blocks every time a weight outgrows it or five updates have passed; 25 iterations cover several
    replacements on a graph with 20 % outlier loop closures."""
    g, _, X0 = D.city2d(14, 12, outlier_fraction=0.2, seed=9)
    out = parity.run_both(g, 4, X0, 25, loss=loss, algorithm=alg, rescale="Dynamic")
    _check(out, 2)
    assert max(out["drv"].node_scalars(a).reserved for a in range(4)) >= 3        # rescales happened
    # and it is a different algorithm than Static: the traces part once the first rescale has happened
    ref = parity.run_both(g, 4, X0, 25, loss=loss, algorithm=alg)
    assert np.abs(out["fobj_gpu"].sum(axis=1) - ref["fobj_gpu"].sum(axis=1)).max() > 1e-6 * out["fobj_gpu"].sum(axis=1)[0]


def test_dynamic_rescale_se3_reinitialize_and_sharded():
    """SE(3), re-initialisation of a handle that has rescaled (the all-ones vector comes back), and a sharded run."""
    from inproc_world import InProcWorld
    g, _, X0 = D.grid3d(8, 8, 6, seed=7)
    out = parity.run_both(g, 4, X0, 14, loss="welsch", algorithm="hash", rescale="Dynamic")
    _check(out, 3)
    drv = out["drv"]
    X1 = drv.X()
    assert drv.initialize(X0) == 0 and drv.update() == 0
    for _ in range(14):
        assert drv.iterate() == 0 and drv.communicate() == 0 and drv.update() == 0
    assert np.array_equal(drv.X(), X1)
    world = InProcWorld(g, 4, 2, "hash", loss="welsch", rescale="Dynamic")
    trace, X = world.run(X0, 14)
    assert np.array_equal(X, X1)


def test_dynamic_rescale_refuses_a_static_factor():
    g, _, _ = D.grid3d(6, 6, 6, seed=1)
    with pytest.raises(D.MmpgoError):
        D.DPGOHash(g, 4, D.Options(loss="huber", rescale="Dynamic", translation_solver="direct"))


@pytest.mark.parametrize("kernel", ["pcg_ring", "pcg_lite"])
def test_pcg_reports_a_solve_that_stops_on_max_iters(kernel):
    """The reference's L_.solve is exact (CHOLMOD, DPGOProblem.cpp:140, 568).  A PCG solve cut off at
    translation_solve_max_iters above translation_solve_tol is reported with MMPGO_ERR_NOT_CONVERGED and counted,
    by the bare solve and by the driver methods; with the default limit the same nodes are solved to scipy's
    answer and nothing is reported."""
    import scipy.sparse.linalg as spla
    from dpgo_b200 import lib as L
    (g, _, X0), nodes = D.sphere_rings(2, 6000, seed=4), 2
    A = _g00(g, nodes)
    rng = np.random.default_rng(9)
    b = rng.standard_normal((g.num_poses, g.d)) * 10.0
    want = -spla.splu(A.tocsc()).solve(b)
    ok = D.DPGOHash(g, nodes, D.Options(dense_solve_max_n=0, translation_solver=kernel))
    t = ok.translation_solve(b)
    assert np.abs(t - want).max() <= 1e-10 * np.abs(want).max()
    c = ok.counters()
    need = c.reserved[4]                                        # most iterations one node took
    assert c.reserved[3] == 0 and 5 < need < 4000
    drv = D.DPGOHash(g, nodes, D.Options(dense_solve_max_n=0, translation_solver=kernel, translation_solve_max_iters=5))
    with pytest.raises(L.MmpgoError) as e:
        drv.translation_solve(b)
    assert e.value.code == -5 and "translation_solve_max_iters" in str(e.value)
    c = drv.counters()
    assert c.reserved[3] == 2 and c.reserved[4] == 5            # both nodes stopped at the limit
    # the driver methods report it as well, once per offending call
    drv.reset_counters()
    assert drv.initialize(X0) == 0
    rcs = [drv.update(), drv.iterate(), drv.update()]
    assert -5 in rcs and all(rc in (0, -5) for rc in rcs)
    assert drv.node_scalars(0).translation_solve_iters == 5


@pytest.mark.parametrize("alg,loss,d", [("hash", "trivial", 3), ("star", "huber", 3), ("hash", "gm", 2)])
def test_regularized_cholesky_preconditioner(alg, loss, d):
    """Preconditioner::RegularizedCholesky, the reference's default (DPGO_types.h:155; DPGOProblem.cpp:101-124,
    592-594): the Cholesky factor of G11 + (lambda_max / 1e6) I applied inside every truncated-CG iteration.  On the
    device: the multifrontal factor of every node's G11 (d scalar rows per pose) and the supernodal sweeps of the
    translation solve.  Same tolerances as the block-Jacobi runs; the largest-eigenvalue estimate is checked against
    scipy's."""
    import scipy.sparse.linalg as spla
    from oracle import dpgo as odpgo
    from oracle import g2o as og2o
    if d == 3:
        g, _, X0 = D.grid3d(7, 6, 5, seed=21)
    else:
        g, _, X0 = D.city2d(14, 10, seed=21)
    nodes = 4
    out = parity.run_both(g, nodes, X0, 12, loss=loss, algorithm=alg, preconditioner="RegularizedCholesky")
    _check(out, d)
    drv = out["drv"]
    c = drv.counters()
    assert c.reserved[5] > 0                                   # the sparse preconditioner solves ran
    assert c.tcg_iterations > 0 or d == 2                      # (the SE(2) case converges without a tCG step)
    per_node, g_index, _ = og2o.partition(g.num_poses, nodes, parity.to_measurements(g))
    for a in range(nodes):
        prob = odpgo.DPGOProblem(a, per_node[a], odpgo.Options(loss=loss, preconditioner="None"))
        want = float(spla.eigsh(prob.G11, k=1, which="LA", tol=1e-10, return_eigenvectors=False)[0])
        lam, nnz = drv.preconditioner_info(a)
        assert abs(lam - want) <= 2e-4 * want and nnz > 0      # the reference's own tolerance is 1e-4
    # the default (block-Jacobi) handle does not carry the factor
    with pytest.raises(D.lib.MmpgoError):
        D.DPGOHash(g, nodes).preconditioner_info(0)


def test_reference_default_options_on_a_reference_dataset():
    """DPGO::Options as the struct declares them (DPGO_types.h:128, 155): Rescale::Dynamic and the
    RegularizedCholesky preconditioner, AMM-PGO* with the Huber loss on the reference's smallGrid3D dataset
    (the graph of the golden fixture), 40 iterations against the oracle run live with the same options."""
    g, z = load_golden("smallGrid3D_n4_huber_star")
    out = parity.run_both(g, 4, z["X0"], 40, loss="huber", algorithm="star", preconditioner="RegularizedCholesky",
                          rescale="Dynamic")
    _check(out, g.d)
    # ... and dist_pgo's own combination (Static rescale, dist_pgo.cpp:105) with the default preconditioner
    out = parity.run_both(g, 4, z["X0"], 40, loss="huber", algorithm="star", preconditioner="RegularizedCholesky")
    _check(out, g.d)
    # sharded handles follow the same trajectory bit for bit (the factor of a node does not depend on its neighbours' GPU)
    from inproc_world import InProcWorld
    world = InProcWorld(g, 4, 2, "star", loss="huber", preconditioner="RegularizedCholesky")
    _, X = world.run(z["X0"], 40)
    assert np.array_equal(X, out["X_gpu"])


@pytest.mark.parametrize("case", ["grid_ragged", "city", "rings"])
def test_set_graph_partition_matches_generate_data_info(case):
    """mmpgo_set_graph takes the flat global edge arrays and partitions them in C++ (the replacement of
    read_g2o's partition + generate_data_info, DPGO_utils.cpp:140-202, 326-438): per node the numbers of own
    poses, neighbour poses, intra- and inter-node measurements, and the poses each GPU has to receive, against
    the oracle's restatement of generate_data_info -- ragged partitions, outlier edges, ring-shaped robots."""
    from oracle import data_matrix as dm, g2o as og2o
    if case == "grid_ragged":
        (g, _, _), nodes = D.grid3d(7, 5, 5, seed=5), 6            # 175 poses over 6 nodes: 30,29,29,29,29,29
    elif case == "city":
        (g, _, _), nodes = D.city2d(23, 17, seed=5), 7
    else:
        (g, _, _), nodes = D.sphere_rings(5, 333, seed=5), 5
    meas = parity.to_measurements(g)
    per_node, g_index, part = og2o.partition(g.num_poses, nodes, meas)
    drv = D.DPGOHash(g, nodes)
    tot_own = 0
    for a in range(nodes):
        info = dm.generate_data_info(a, per_node[a])
        s = drv.node_scalars(a)
        assert (s.n0, s.n1, s.m0, s.m1) == (info.n[0], info.n[1], info.m[0], info.m[1]), (a, case)
        tot_own += s.n0
    sz = drv.sizes()
    assert tot_own == g.num_poses == sz["own_poses"] and sz["halo_poses"] == 0 and sz["local_nodes"] == nodes
    # a sharded handle owns the nodes of its rank and lists exactly the remote neighbour poses as its halo
    half = nodes // 2
    lo = D.DPGOHash(g, nodes, None, 0, half)
    want = set()
    for a in range(half):
        info = dm.generate_data_info(a, per_node[a])
        for k in info.nbr_keys:
            if int(k >> 40) >= half:
                want.add(int(k))
    assert lo.sizes()["halo_poses"] == len(want) and lo.sizes()["local_nodes"] == half
