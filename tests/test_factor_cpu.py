"""The sparse direct solver behind the translation solve (host side: nested dissection,
multifrontal Cholesky, the host restatement of the device sweeps) against scipy, on the matrices
the hot path factors: G00 of a robot node (tau-weighted intra-node Laplacian + 2 tau per
inter-node edge + xi, DPGO_utils.cpp:2212-2243) and block matrices like G11 + lambda I."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import dpgo_b200 as D
from dpgo_b200 import lib as L


def mf_solve(A, rhs, block=1, leaf=16):
    A = sp.csr_matrix(A)
    A.sort_indices()
    n, nrhs = A.shape[0], rhs.shape[1]
    ptr, col = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    val = np.ascontiguousarray(A.data, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x = np.zeros_like(rhs)
    stats = np.zeros(8, dtype=np.int64)
    L.check(L.load().mmpgo_mf_host_solve(n, L.iptr(ptr), L.iptr(col), L.dptr(val), block, leaf, nrhs, L.dptr(rhs),
                                         L.dptr(x), stats.ctypes.data_as(C.POINTER(C.c_int64))))
    return x, stats


def g00_of_node(g, num_nodes, a, xi=1e-11):
    n0 = g.num_poses // num_nodes
    lo, hi = a * n0, (a + 1) * n0
    ii, jj = g.i.astype(np.int64), g.j.astype(np.int64)
    own_i, own_j = (ii >= lo) & (ii < hi), (jj >= lo) & (jj < hi)
    intra = own_i & own_j
    A = sp.coo_matrix((np.concatenate([-g.tau[intra]] * 2),
                       (np.concatenate([ii[intra], jj[intra]]) - lo, np.concatenate([jj[intra], ii[intra]]) - lo)),
                      shape=(n0, n0)).tocsr()
    diag = np.full(n0, xi)
    np.add.at(diag, ii[intra] - lo, g.tau[intra])
    np.add.at(diag, jj[intra] - lo, g.tau[intra])
    np.add.at(diag, ii[own_i & ~own_j] - lo, 2 * g.tau[own_i & ~own_j])
    np.add.at(diag, jj[own_j & ~own_i] - lo, 2 * g.tau[own_j & ~own_i])
    return (A + sp.diags(diag)).tocsr()


@pytest.mark.parametrize("dims,nodes,a", [((20, 20, 12), 3, 1), ((100, 125, 5), 4, 2), ((9, 7, 5), 1, 0)])
@pytest.mark.parametrize("leaf", [4, 16, 64])
def test_g00_solve_matches_scipy(dims, nodes, a, leaf):
    g, _, _ = D.grid3d(*dims, seed=3)
    A = g00_of_node(g, nodes, a, xi=1e-11 if nodes > 1 else 1e-3)
    rng = np.random.default_rng(0)
    b = rng.standard_normal((A.shape[0], 3))
    x, stats = mf_solve(A, b, leaf=leaf)
    want = spla.spsolve(A.tocsc(), b)
    # the single-node matrix is a Laplacian + 1e-3 I (condition number ~1e7): forward error scales with it
    assert np.abs(x - want).max() <= (1e-11 if nodes > 1 else 1e-7) * np.abs(want).max()
    assert np.abs(A @ x - b).max() <= 1e-10 * np.abs(b).max()
    assert stats[0] >= A.shape[0] and stats[1] >= 0


def test_chain_and_ring_nodes():
    """Odometry chains and rings (the multi-robot sphere of BASELINE.json configs[4]): separators of
    one or two poses, a tree of logarithmic height."""
    n = 5000
    rng = np.random.default_rng(1)
    w = rng.uniform(100, 500, n)
    i = np.arange(n)
    j = (i + 1) % n
    A = sp.coo_matrix((np.concatenate([-w, -w]), (np.concatenate([i, j]), np.concatenate([j, i]))), shape=(n, n)).tocsr()
    A = A + sp.diags(-np.asarray(A.sum(axis=1)).ravel() + rng.uniform(50, 100, n))
    b = rng.standard_normal((n, 3))
    x, stats = mf_solve(A, b)
    assert np.abs(A @ x - b).max() <= 1e-9 * np.abs(b).max()
    assert stats[1] <= 2 * int(np.ceil(np.log2(n)))          # tree height
    assert stats[0] <= 12 * n                                   # nnz(L) stays linear


def test_block_matrix_and_duplicates():
    """block > 1 (d rows per pose, the shape of G11 + lambda I) and duplicate CSR entries
    (parallel edges between the same two poses are summed)."""
    g, _, _ = D.grid3d(8, 8, 6, seed=5)
    n, d = g.num_poses, 3
    rng = np.random.default_rng(2)
    rows, cols, vals = [], [], []
    for e in range(g.num_edges):
        B = rng.standard_normal((d, d))
        for r in range(d):
            for c in range(d):
                rows += [d * g.i[e] + r, d * g.j[e] + c]
                cols += [d * g.j[e] + c, d * g.i[e] + r]
                vals += [B[r, c], B[r, c]]
    A = sp.coo_matrix((vals, (rows, cols)), shape=(d * n, d * n)).tocsr()
    A = A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)
    # duplicates: split every entry into two halves stored separately
    coo = A.tocoo()
    ptr = np.zeros(d * n + 1, dtype=np.int32)
    order = np.lexsort((np.tile(np.arange(2), len(coo.data)), np.repeat(coo.col, 2), np.repeat(coo.row, 2)))
    r2, c2, v2 = np.repeat(coo.row, 2)[order], np.repeat(coo.col, 2)[order], np.repeat(coo.data / 2, 2)[order]
    np.add.at(ptr, r2 + 1, 1)
    ptr = np.cumsum(ptr).astype(np.int32)
    b = rng.standard_normal((d * n, d))
    x = np.zeros_like(b)
    col = c2.astype(np.int32)
    L.check(L.load().mmpgo_mf_host_solve(d * n, L.iptr(ptr), L.iptr(col), L.dptr(np.ascontiguousarray(v2)), d, 16, d,
                                         L.dptr(b), L.dptr(x), None))
    assert np.abs(A @ x - b).max() <= 1e-10 * np.abs(b).max()


def test_not_positive_definite_is_reported():
    A = sp.csr_matrix(np.array([[1.0, 2.0], [2.0, 1.0]]))
    with pytest.raises(L.MmpgoError):
        mf_solve(A, np.ones((2, 1)))
