"""Pose-graph containers, g2o reader and synthetic generators (host side).

`read_g2o` mirrors DPGO::read_g2o_file (C++/DPGO/src/DPGO_utils.cpp:8-138):
EDGE_SE2 / EDGE_SE3:QUAT lines, tau = d / tr(I_t^{-1}), kappa = I_33 (2-D) or
3 / (2 tr(I_R^{-1})) (3-D); VERTEX lines are ignored.  The generators build the
synthetic graphs named in BASELINE.json (3-D grid, multi-robot sphere, 2-D
city with injected outliers) as the same edge arrays, so they go through the
same partition and upload path as a g2o file.
"""
from __future__ import annotations

import numpy as np


class PoseGraph:
    """Global edge list: edge k measures pose j[k] in the frame of pose i[k]."""

    def __init__(self, d, num_poses, i, j, R, t, kappa, tau):
        self.d = int(d)
        self.num_poses = int(num_poses)
        self.i = np.ascontiguousarray(i, dtype=np.int32)
        self.j = np.ascontiguousarray(j, dtype=np.int32)
        self.R = np.ascontiguousarray(R, dtype=np.float64).reshape(-1, d, d)
        self.t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1, d)
        self.kappa = np.ascontiguousarray(kappa, dtype=np.float64)
        self.tau = np.ascontiguousarray(tau, dtype=np.float64)

    @property
    def num_edges(self):
        return len(self.kappa)


def _quat_to_rot(qw, qx, qy, qz):
    # Eigen::Quaternion::toRotationMatrix, no normalisation (DPGO_utils.cpp:100-101)
    R = np.empty(qw.shape + (3, 3))
    tx, ty, tz = 2 * qx, 2 * qy, 2 * qz
    R[..., 0, 0] = 1 - (ty * qy + tz * qz)
    R[..., 0, 1] = ty * qx - tz * qw
    R[..., 0, 2] = tz * qx + ty * qw
    R[..., 1, 0] = ty * qx + tz * qw
    R[..., 1, 1] = 1 - (tx * qx + tz * qz)
    R[..., 1, 2] = tz * qy - tx * qw
    R[..., 2, 0] = tz * qx - ty * qw
    R[..., 2, 1] = tz * qy + tx * qw
    R[..., 2, 2] = 1 - (tx * qx + ty * qy)
    return R


def read_g2o(path):
    rows2, rows3 = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("EDGE_SE3:QUAT"):
                rows3.append(line.split()[1:31])
            elif line.startswith("EDGE_SE2"):
                rows2.append(line.split()[1:12])
            elif line.startswith("VERTEX") or not line.strip():
                continue
            else:
                raise ValueError("unrecognized g2o record: %s" % line.split()[0])
    if rows3 and rows2:
        raise ValueError("mixed SE(2)/SE(3) file")
    if rows3:
        a = np.array(rows3, dtype=np.float64)
        i, j = a[:, 0].astype(np.int32), a[:, 1].astype(np.int32)
        t = a[:, 2:5]
        R = _quat_to_rot(a[:, 8], a[:, 5], a[:, 6], a[:, 7])
        I = a[:, 9:]
        It = np.stack([I[:, [0, 1, 2]], I[:, [1, 6, 7]], I[:, [2, 7, 11]]], axis=1)
        Ir = np.stack([I[:, [15, 16, 17]], I[:, [16, 18, 19]], I[:, [17, 19, 20]]], axis=1)
        tau = 3.0 / np.trace(np.linalg.inv(It), axis1=1, axis2=2)
        kappa = 3.0 / (2.0 * np.trace(np.linalg.inv(Ir), axis1=1, axis2=2))
        d = 3
    else:
        a = np.array(rows2, dtype=np.float64)
        i, j = a[:, 0].astype(np.int32), a[:, 1].astype(np.int32)
        t = a[:, 2:4]
        c, s = np.cos(a[:, 4]), np.sin(a[:, 4])
        R = np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], axis=1)
        It = np.stack([a[:, [5, 6]], a[:, [6, 8]]], axis=1)
        tau = 2.0 / np.trace(np.linalg.inv(It), axis1=1, axis2=2)
        kappa = a[:, 10].copy()
        d = 2
    n = int(max(i.max(), j.max())) + 1
    return PoseGraph(d, n, i, j, R, t, kappa, tau)


def _rot_to_quat(R):
    # (qx, qy, qz, qw) of a rotation matrix, branch on the largest diagonal term
    q = np.empty((len(R), 4))
    for k, M in enumerate(R):
        tr = M[0, 0] + M[1, 1] + M[2, 2]
        if tr > 0:
            s = 2.0 * np.sqrt(tr + 1.0)
            q[k] = ((M[2, 1] - M[1, 2]) / s, (M[0, 2] - M[2, 0]) / s, (M[1, 0] - M[0, 1]) / s, 0.25 * s)
        else:
            i = int(np.argmax(np.diag(M)))
            j, l = (i + 1) % 3, (i + 2) % 3
            s = 2.0 * np.sqrt(1.0 + M[i, i] - M[j, j] - M[l, l])
            v = np.empty(4)
            v[i] = 0.25 * s
            v[j] = (M[j, i] + M[i, j]) / s
            v[l] = (M[l, i] + M[i, l]) / s
            v[3] = (M[l, j] - M[j, l]) / s
            q[k] = v
    return q


def write_g2o(path, g):
    """Writes a PoseGraph as EDGE_SE2 / EDGE_SE3:QUAT lines with isotropic information matrices
    chosen so that the reader's formulas (DPGO_utils.cpp:63-67, 107-116) give back tau and kappa."""
    with open(path, "w") as fh:
        if g.d == 3:
            q = _rot_to_quat(g.R)
            for k in range(g.num_edges):
                it, ir = g.tau[k], 2.0 * g.kappa[k]
                info = [it, 0, 0, 0, 0, 0, it, 0, 0, 0, 0, it, 0, 0, 0, ir, 0, 0, ir, 0, ir]
                fh.write("EDGE_SE3:QUAT %d %d %s %s %s\n" % (
                    g.i[k], g.j[k], " ".join(repr(float(v)) for v in g.t[k]),
                    " ".join(repr(float(v)) for v in q[k]), " ".join(repr(float(v)) for v in info)))
        else:
            for k in range(g.num_edges):
                th = np.arctan2(g.R[k, 1, 0], g.R[k, 0, 0])
                it = g.tau[k]
                fh.write("EDGE_SE2 %d %d %r %r %r %r 0.0 0.0 %r 0.0 %r\n" % (
                    g.i[k], g.j[k], float(g.t[k, 0]), float(g.t[k, 1]), float(th), float(it), float(it),
                    float(g.kappa[k])))


# ---------------------------------------------------------------------------
# synthetic graphs
# ---------------------------------------------------------------------------
def so3_exp(w):
    th = np.linalg.norm(w, axis=-1)
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    a = np.where(small, 1.0 - th * th / 6.0, np.sin(ths) / ths)
    b = np.where(small, 0.5 - th * th / 24.0, (1.0 - np.cos(ths)) / (ths * ths))
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    return np.eye(3) + a[..., None, None] * K + b[..., None, None] * (K @ K)


def so2_exp(th):
    c, s = np.cos(th), np.sin(th)
    return np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], axis=-2)


def _measure(rng, Rg, tg, i, j, sig_t, sig_r, d):
    """Noisy relative measurements of the ground truth (Rg, tg) on edges i->j."""
    Ri = Rg[i]
    RiT = np.swapaxes(Ri, 1, 2)
    t = np.einsum("eab,eb->ea", RiT, tg[j] - tg[i]) + sig_t * rng.standard_normal((len(i), d))
    if d == 3:
        N = so3_exp(sig_r * rng.standard_normal((len(i), 3)))
    else:
        N = so2_exp(sig_r * rng.standard_normal(len(i)))
    R = RiT @ Rg[j] @ N
    return R, t


def to_global_X(Rg, tg):
    """Reference global layout [t (N rows); R_i^T blocks (dN rows)]."""
    N, d = tg.shape
    return np.vstack([tg, np.swapaxes(Rg, 1, 2).reshape(N * d, d)])


def perturbed_init(rng, Rg, tg, sig_t, sig_r):
    N, d = tg.shape
    if d == 3:
        Rn = Rg @ so3_exp(sig_r * rng.standard_normal((N, 3)))
    else:
        Rn = Rg @ so2_exp(sig_r * rng.standard_normal(N))
    return to_global_X(Rn, tg + sig_t * rng.standard_normal((N, d)))


def grid3d(nx, ny, nz, num_edges=None, seed=20241017, sig_t=0.05, sig_r=0.03,
           init_sig_t=0.2, init_sig_r=0.1):
    """SE(3) lattice of nx*ny*nz poses in serpentine ("snake") id order, so that
    consecutive ids are lattice neighbours and contiguous id ranges are slabs.
    Edges: the odometry path, all remaining lattice-neighbour pairs, and in-plane
    diagonals sampled to reach `num_edges` (default 4 per pose).  Information
    matrix diag(1/sig_t^2 x3, 1/sig_r^2 x3) => tau = 1/sig_t^2, kappa = 1/(2 sig_r^2)
    through the reader's formulas (DPGO_utils.cpp:107-116).
    Returns (PoseGraph, X_ground_truth, X_init)."""
    rng = np.random.default_rng(seed)
    N = nx * ny * nz
    if num_edges is None:
        num_edges = 4 * N
    z, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    y = np.where(z % 2 == 0, yy, ny - 1 - yy)
    row = z * ny + yy
    x = np.where(row % 2 == 0, xx, nx - 1 - xx)
    ids = np.arange(N).reshape(nz, ny, nx)
    coord_to_id = np.empty((nz, ny, nx), dtype=np.int64)
    coord_to_id[z, y, x] = ids
    tg = np.empty((N, 3))
    tg[coord_to_id.ravel()] = np.stack(np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx),
                                                   indexing="ij")[::-1], -1).reshape(-1, 3)
    Rg = so3_exp(0.3 * rng.standard_normal((N, 3)))
    c = coord_to_id
    pairs = [np.stack([c[:, :, :-1].ravel(), c[:, :, 1:].ravel()], 1),
             np.stack([c[:, :-1, :].ravel(), c[:, 1:, :].ravel()], 1),
             np.stack([c[:-1, :, :].ravel(), c[1:, :, :].ravel()], 1)]
    lat = np.concatenate(pairs)
    lat = np.stack([lat.min(1), lat.max(1)], 1)
    odo = lat[:, 1] == lat[:, 0] + 1
    e = [np.stack([np.arange(N - 1), np.arange(1, N)], 1), lat[~odo]]
    have = N - 1 + int((~odo).sum())
    if have > num_edges:
        keep = rng.permutation(int((~odo).sum()))[: num_edges - (N - 1)]
        e[1] = lat[~odo][np.sort(keep)]
    elif have < num_edges:
        diag = np.concatenate([
            np.stack([c[:, :-1, :-1].ravel(), c[:, 1:, 1:].ravel()], 1),
            np.stack([c[:, :-1, 1:].ravel(), c[:, 1:, :-1].ravel()], 1)])
        need = min(num_edges - have, len(diag))
        sel = np.sort(rng.permutation(len(diag))[:need])
        dsel = diag[sel]
        e.append(np.stack([dsel.min(1), dsel.max(1)], 1))
    E = np.concatenate(e)
    i, j = E[:, 0], E[:, 1]
    R, t = _measure(rng, Rg, tg, i, j, sig_t, sig_r, 3)
    m = len(i)
    g = PoseGraph(3, N, i, j, R, t, np.full(m, 1.0 / (2.0 * sig_r ** 2)), np.full(m, 1.0 / sig_t ** 2))
    return g, to_global_X(Rg, tg), perturbed_init(rng, Rg, tg, init_sig_t, init_sig_r)


def sphere_rings(num_robots, poses_per_robot, seed=20241018, sig_t=0.05, sig_r=0.03,
                 init_sig_t=0.2, init_sig_r=0.1, radius=50.0):
    """Multi-robot sphere: robot r drives one latitude ring; intra-robot odometry
    plus ring closure, and closures to the next ring at the same / next /
    previous longitude index (about 4 edges per pose)."""
    rng = np.random.default_rng(seed)
    P, n = num_robots, poses_per_robot
    N = P * n
    r = np.repeat(np.arange(P), n)
    k = np.tile(np.arange(n), P)
    lat = (r + 1) / (P + 1) * np.pi - np.pi / 2
    lon = 2 * np.pi * k / n
    tg = radius * np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], -1)
    Rg = so3_exp(np.stack([np.zeros(N), np.zeros(N), lon + np.pi / 2], -1)) @ \
        so3_exp(0.1 * rng.standard_normal((N, 3)))
    ids = np.arange(N).reshape(P, n)
    e = [np.stack([ids[:, :-1].ravel(), ids[:, 1:].ravel()], 1),
         np.stack([ids[:, 0], ids[:, -1]], 1)]
    for sh in (0, 1, -1):
        e.append(np.stack([ids[:-1].ravel(), np.roll(ids[1:], -sh, axis=1).ravel()], 1))
    E = np.concatenate(e)
    i, j = E.min(1), E.max(1)
    R, t = _measure(rng, Rg, tg, i, j, sig_t, sig_r, 3)
    m = len(i)
    g = PoseGraph(3, N, i, j, R, t, np.full(m, 1.0 / (2.0 * sig_r ** 2)), np.full(m, 1.0 / sig_t ** 2))
    return g, to_global_X(Rg, tg), perturbed_init(rng, Rg, tg, init_sig_t, init_sig_r)


def city2d(nx, ny, outlier_fraction=0.1, seed=20241019, sig_t=0.1, sig_r=0.02,
           init_sig_t=0.3, init_sig_r=0.1):
    """SE(2) Manhattan-style grid path (snake order) with loop closures between
    adjacent rows; `outlier_fraction` of the loop closures are replaced by
    uniformly random measurements (the injected outliers of BASELINE.json
    config 3)."""
    rng = np.random.default_rng(seed)
    N = nx * ny
    yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    x = np.where(yy % 2 == 0, xx, nx - 1 - xx)
    c = np.empty((ny, nx), dtype=np.int64)
    c[yy, x] = np.arange(N).reshape(ny, nx)
    tg = np.empty((N, 2))
    tg[c.ravel()] = np.stack([xx, yy], -1).reshape(-1, 2).astype(float)
    Rg = so2_exp(0.5 * rng.standard_normal(N))
    lc = np.stack([c[:-1, :].ravel(), c[1:, :].ravel()], 1)
    lc = np.stack([lc.min(1), lc.max(1)], 1)
    lc = lc[lc[:, 1] != lc[:, 0] + 1]
    E = np.concatenate([np.stack([np.arange(N - 1), np.arange(1, N)], 1), lc])
    i, j = E[:, 0], E[:, 1]
    R, t = _measure(rng, Rg, tg, i, j, sig_t, sig_r, 2)
    n_lc = len(lc)
    n_out = int(round(outlier_fraction * n_lc))
    out = (N - 1) + np.sort(rng.permutation(n_lc)[:n_out])
    R[out] = so2_exp(rng.uniform(-np.pi, np.pi, n_out))
    t[out] = rng.uniform(-5.0, 5.0, (n_out, 2))
    m = len(i)
    g = PoseGraph(2, N, i, j, R, t, np.full(m, 1.0 / sig_r ** 2), np.full(m, 1.0 / sig_t ** 2))
    g.outlier_edges = out
    return g, to_global_X(Rg, tg), perturbed_init(rng, Rg, tg, init_sig_t, init_sig_r)
