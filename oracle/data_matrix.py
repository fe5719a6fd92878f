"""Per-node data matrices (oracle; test infrastructure only).

Restates generate_data_info (C++/DPGO/src/DPGO_utils.cpp:326-438),
simplify_quadratic_data_matrix (:1398-2288, trivial loss) and the Static
simplify_regular_data_matrix (:2290-2967, robust losses) by assembling the
same scalar triplets into scipy CSR matrices.

Row/column numbering of the stacked variable Z = [t_own; R_own; t_nbr; R_nbr]
(DPGO_utils.cpp:1524-1531): own translation i -> i, own rotation row k of
pose i -> n0 + d*i + k, neighbour translation j -> (d+1)*n0 + j, neighbour
rotation row k of pose j -> (d+1)*n0 + n1 + d*j + k.

Per-edge blocks (rows/cols ordered [t, Y_0..Y_{d-1}]), from the triplets at
DPGO_utils.cpp:1542-1641:
    Mii = [[tau, tau t^T], [tau t, kappa I + tau t t^T]]
    Mjj = [[tau, 0], [0, kappa I]]
    Mij = [[-tau, 0], [-tau t, -kappa R]],  Mji = Mij^T
so that M_e = b_e^T b_e for the (d+1) rows b_e of B (DPGO_utils.cpp:1643-1676):
    row 0   : sqrt(tau)   (t_i - t_j + t_e^T Y_i)
    row 1+r : sqrt(kappa) (R_e^T Y_i - Y_j)[r, :]
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class DataInfo:
    pass


def generate_data_info(a, meas):
    """DPGO_utils.cpp:326-438 (vectorised; same ordering rules: own poses by
    ascending pose id, neighbour poses by (node, pose), :400-409)."""
    info = DataInfo()
    intra_mask = (meas.i_node == a) & (meas.j_node == a)
    info.intra = meas.select(intra_mask)
    info.inter = meas.select(~intra_mask)
    BIG = np.int64(1) << 40
    nodes = np.concatenate([meas.i_node, meas.j_node])
    poses = np.concatenate([meas.i_pose, meas.j_pose])
    other_nodes = np.concatenate([meas.j_node, meas.i_node])
    own_mask = nodes == a
    info.own_poses = np.unique(poses[own_mask])
    nbr_keys = np.unique(nodes[~own_mask] * BIG + poses[~own_mask])
    info.nbr_keys = nbr_keys
    info.nbr_poses = [(int(k >> 40), int(k & (BIG - 1))) for k in nbr_keys] \
        if len(nbr_keys) < 200000 else None
    info.n = (len(info.own_poses), len(nbr_keys))
    info.m = (len(info.intra), len(info.inter))
    # sent[b]: own poses with a measurement to node b (:384-392, :426-433)
    sel = own_mask & (other_nodes != a)
    sk = np.unique(other_nodes[sel] * BIG + poses[sel])
    sent = {}
    for k in sk:
        sent.setdefault(int(k >> 40), []).append(int(k & (BIG - 1)))
    info.sent = sent
    recv = {}
    for k in nbr_keys:
        recv.setdefault(int(k >> 40), []).append(int(k & (BIG - 1)))
    info.recv = recv
    info.d = meas.d
    info.node = a
    return info


def _edge_blocks(meas):
    d = meas.d
    m = len(meas)
    tau, kap, t, R = meas.tau, meas.kappa, meas.t, meas.R
    Mii = np.zeros((m, d + 1, d + 1))
    Mjj = np.zeros((m, d + 1, d + 1))
    Mij = np.zeros((m, d + 1, d + 1))
    Mii[:, 0, 0] = tau
    Mii[:, 0, 1:] = tau[:, None] * t
    Mii[:, 1:, 0] = tau[:, None] * t
    Mii[:, 1:, 1:] = (tau[:, None, None] * t[:, :, None]) * t[:, None, :]
    for k in range(d):
        Mii[:, 1 + k, 1 + k] += kap
        Mjj[:, 1 + k, 1 + k] = kap
    Mjj[:, 0, 0] = tau
    Mij[:, 0, 0] = -tau
    Mij[:, 1:, 0] = -tau[:, None] * t
    Mij[:, 1:, 1:] = -kap[:, None, None] * R
    return Mii, Mjj, Mij


def _pose_rows(info, node, pose):
    """Z-row indices [t, Y_0..Y_{d-1}] of poses given as (node, pose) arrays."""
    d, (n0, n1), a = info.d, info.n, info.node
    node = np.asarray(node, dtype=np.int64)
    pose = np.asarray(pose, dtype=np.int64)
    rows = np.empty((len(node), d + 1), dtype=np.int64)
    own = node == a
    lo = np.searchsorted(info.own_poses, pose[own])
    ln = np.searchsorted(info.nbr_keys, node[~own] * (np.int64(1) << 40) + pose[~own])
    ar = np.arange(d)
    rows[own, 0] = lo
    rows[own, 1:] = n0 + d * lo[:, None] + ar
    rows[~own, 0] = (d + 1) * n0 + ln
    rows[~own, 1:] = (d + 1) * n0 + n1 + d * ln[:, None] + ar
    return rows


def local_index(info, node, pose):
    """(is_own, local index) of poses given as (node, pose) arrays."""
    node = np.asarray(node, dtype=np.int64)
    pose = np.asarray(pose, dtype=np.int64)
    own = node == info.node
    idx = np.empty(len(node), dtype=np.int64)
    idx[own] = np.searchsorted(info.own_poses, pose[own])
    idx[~own] = np.searchsorted(info.nbr_keys,
                                node[~own] * (np.int64(1) << 40) + pose[~own])
    return own, idx


class _Trip:
    def __init__(self):
        self.r, self.c, self.v = [], [], []

    def block(self, rows, cols, blk, scale=1.0, mask=None):
        """rows, cols: (m, d+1); blk: (m, d+1, d+1)."""
        if mask is not None:
            rows, cols, blk = rows[mask], cols[mask], blk[mask]
        if len(rows) == 0:
            return
        k = rows.shape[1]
        self.r.append(np.repeat(rows, k, axis=1).ravel())
        self.c.append(np.tile(cols, (1, k)).ravel())
        self.v.append((scale * blk).ravel())

    def diag(self, idx, val):
        idx = np.asarray(idx, dtype=np.int64)
        self.r.append(idx); self.c.append(idx)
        self.v.append(np.full(len(idx), float(val)))

    def csr(self, shape):
        if not self.r:
            return sp.csr_matrix(shape)
        r, c, v = (np.concatenate(x) for x in (self.r, self.c, self.v))
        nz = v != 0.0
        return sp.coo_matrix((v[nz], (r[nz], c[nz])), shape=shape).tocsr()


def _B_rows(meas, I, J, ncols):
    """Rows of the incidence matrix B (DPGO_utils.cpp:1643-1676)."""
    d, m = meas.d, len(meas)
    st, sk = np.sqrt(meas.tau), np.sqrt(meas.kappa)
    r, c, v = [], [], []
    l = (d + 1) * np.arange(m)
    r += [l, l]; c += [I[:, 0], J[:, 0]]; v += [st, -st]
    for k in range(d):
        r.append(l); c.append(I[:, 1 + k]); v.append(st * meas.t[:, k])
    for rr in range(d):
        for cc in range(d):
            r.append(l + rr + 1); c.append(I[:, 1 + cc])
            v.append(sk * meas.R[:, cc, rr])
        r.append(l + rr + 1); c.append(J[:, 1 + rr]); v.append(-sk)
    if m == 0:
        return sp.csr_matrix(((d + 1) * m, ncols))
    r, c, v = (np.concatenate(x) for x in (r, c, v))
    return sp.coo_matrix((v, (r, c)), shape=((d + 1) * m, ncols)).tocsr()


def build_data_matrices(info, xi, quadratic, rescale=None, dynamic=False):
    """quadratic=True : simplify_quadratic_data_matrix (DPGO_utils.cpp:1398-2288)
    quadratic=False: simplify_regular_data_matrix, Static (:2290-2967).
    dynamic=True (robust losses only): the Rescale::Dynamic builder (:2969-3903) evaluated at the per-measurement
    rescale vector `rescale` (DiagReScale_, one entry per inter-node measurement; ones at construction,
    DPGOProblem.cpp:31, 84): the inter-node blocks E s / F s that update_quadratic_mat (DPGOProblem.cpp:751-840)
    adds to M, D0, Q0, T0, N0, V0 are the Static blocks scaled by s_e (own endpoint: :3518-3550 E; neighbour
    endpoint: F, enters Q only), and the auxiliary matrices carry 0.5 xi instead of 1.5 xi (:3621, :3635).
    Returns a dict of scipy CSR matrices named as in the reference."""
    d, (n0, n1) = info.d, info.n
    a = info.node
    NX, NZ = (d + 1) * n0, (d + 1) * (n0 + n1)
    intra, inter = info.intra, info.inter

    G, D, S, Q, P, P0, H, V = (_Trip() for _ in range(8))

    # ---- intra-node measurements (:1497-1755 / :2380-2610)
    I = _pose_rows(info, intra.i_node, intra.i_pose)
    J = _pose_rows(info, intra.j_node, intra.j_pose)
    Mii, Mjj, Mij = _edge_blocks(intra)
    Mji = np.transpose(Mij, (0, 2, 1))
    for T_, s in ((G, 1.0), (P, -1.0)) if quadratic else ((G, 1.0),):
        T_.block(I, I, Mii, s); T_.block(J, J, Mjj, s)
        T_.block(I, J, Mij, s); T_.block(J, I, Mji, s)
    H.block(I, I, Mii, 2.0); H.block(J, J, Mjj, 2.0)
    if quadratic:
        V.block(I, I, Mii, -1.0); V.block(J, J, Mjj, -1.0)
        V.block(I, J, Mij, 1.0); V.block(J, I, Mji, 1.0)
    B0 = _B_rows(intra, I, J, NZ)

    # ---- inter-node measurements (:1757-2210 / :2612-2903)
    I = _pose_rows(info, inter.i_node, inter.i_pose)
    J = _pose_rows(info, inter.j_node, inter.j_pose)
    Mii, Mjj, Mij = _edge_blocks(inter)
    Mji = np.transpose(Mij, (0, 2, 1))
    own_i = inter.i_node == a
    own_j = ~own_i
    if dynamic:
        assert not quadratic
        sc = np.ones(len(inter)) if rescale is None else np.asarray(rescale, dtype=float)
        Mii = Mii * sc[:, None, None]
        Mjj = Mjj * sc[:, None, None]
    if quadratic:
        for T_, sd, so in ((Q, -0.5, 0.5), (P0, 0.5, -0.5)):
            # Q = 1/2 M_e - blockdiag ; P0 = blockdiag - 1/2 M_e
            T_.block(I, I, Mii, sd); T_.block(J, J, Mjj, sd)
            T_.block(I, J, Mij, so); T_.block(J, I, Mji, so)
        P.block(I, J, Mij, -1.0); P.block(J, I, Mji, -1.0)
        for T_ in (S, V):
            T_.block(I, I, Mii, -1.0, own_i); T_.block(I, J, Mij, 1.0, own_i)
            T_.block(J, J, Mjj, -1.0, own_j); T_.block(J, I, Mji, 1.0, own_j)
    else:
        Q.block(I, I, Mii, 2.0); Q.block(J, J, Mjj, 2.0)
    for T_ in (G, D, H):
        T_.block(I, I, Mii, 2.0, own_i)
        T_.block(J, J, Mjj, 2.0, own_j)
    B1 = _B_rows(inter, I, J, NZ)

    # ---- regulariser xi (:2212-2243 / :2906-2927)
    own_rows = np.arange(NX)
    G.diag(own_rows, xi); D.diag(own_rows, xi); H.diag(own_rows, (0.5 if dynamic else 1.5) * xi)
    if quadratic:
        S.diag(own_rows, -xi); Q.diag(own_rows, -xi)
        P.diag(own_rows, xi); P0.diag(own_rows, xi)
        V.diag(own_rows, -1.5 * xi)
    else:
        Q.diag(own_rows, 2.0 * xi)

    out = {}
    out["G"] = G.csr((NX, NX))
    out["D"] = D.csr((NX, NX))
    out["Q"] = Q.csr((NZ, NZ))
    out["B0"], out["B1"] = B0, B1
    Hm = H.csr((NX, NX))
    out["H"] = Hm
    out["G00"] = out["G"][:n0, :n0].tocsr()
    out["G01"] = out["G"][:n0, n0:].tocsr()
    out["G10"] = out["G01"].T.tocsr()
    out["G11"] = out["G"][n0:, n0:].tocsr()
    Tdiag = 1.0 / Hm.diagonal()[:n0]          # T = T.inverse(), :2280 / :2958
    out["T"] = Tdiag
    N = sp.diags(Tdiag) @ Hm[:n0, n0:]        # N = T * N, :2282 / :2960
    out["N"] = N.tocsr()
    if quadratic:
        out["S"] = S.csr((NX, NZ))
        out["P"] = P.csr((NZ, NZ))
        out["P0"] = P0.csr((NZ, NZ))
        Vm = V.csr((NX, NZ))
        out["Vfull"] = Vm
        # U = N^T V_top - V_bottom, :2284-2285
        out["U"] = (out["N"].T @ Vm[:n0, :] - Vm[n0:, :]).tocsr()
        # the pose-local V' (used for cross-checks only)
    # V' = H_RR - H_Rt T H_tR, :2962-2964
    K = sp.diags(Tdiag) @ Hm[:n0, n0:]
    out["V"] = (Hm[n0:, n0:] - Hm[n0:, :n0] @ K).tocsr()
    return out
