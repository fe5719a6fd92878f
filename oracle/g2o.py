"""g2o reader + node partition (oracle; test infrastructure only).

Restates C++/DPGO/src/DPGO_utils.cpp:8-138 (read_g2o_file) and :140-202
(read_g2o: contiguous id-range partition, inter-node edges stored in both
endpoint nodes' lists).
"""
from __future__ import annotations

import numpy as np


class Measurements:
    """SoA container of relative pose measurements
    (C++/DPGO/include/DPGO/RelativePoseMeasurement.h:11-29)."""

    def __init__(self, d, i_node, i_pose, j_node, j_pose, R, t, kappa, tau):
        self.d = int(d)
        self.i_node = np.asarray(i_node, dtype=np.int64)
        self.i_pose = np.asarray(i_pose, dtype=np.int64)
        self.j_node = np.asarray(j_node, dtype=np.int64)
        self.j_pose = np.asarray(j_pose, dtype=np.int64)
        self.R = np.asarray(R, dtype=np.float64).reshape(-1, d, d)
        self.t = np.asarray(t, dtype=np.float64).reshape(-1, d)
        self.kappa = np.asarray(kappa, dtype=np.float64)
        self.tau = np.asarray(tau, dtype=np.float64)

    def __len__(self):
        return len(self.kappa)

    def select(self, mask):
        return Measurements(self.d, self.i_node[mask], self.i_pose[mask],
                            self.j_node[mask], self.j_pose[mask], self.R[mask],
                            self.t[mask], self.kappa[mask], self.tau[mask])


def quat_to_rot(qw, qx, qy, qz):
    """Eigen::Quaternion::toRotationMatrix (no normalisation), used at
    DPGO_utils.cpp:100-101."""
    tx, ty, tz = 2 * qx, 2 * qy, 2 * qz
    twx, twy, twz = tx * qw, ty * qw, tz * qw
    txx, txy, txz = tx * qx, ty * qx, tz * qx
    tyy, tyz, tzz = ty * qy, tz * qy, tz * qz
    R = np.empty(np.shape(qw) + (3, 3))
    R[..., 0, 0] = 1 - (tyy + tzz)
    R[..., 0, 1] = txy - twz
    R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz
    R[..., 1, 1] = 1 - (txx + tzz)
    R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy
    R[..., 2, 1] = tyz + twx
    R[..., 2, 2] = 1 - (txx + tyy)
    return R


def read_g2o_file(filename):
    """DPGO_utils.cpp:8-138.  Returns (num_poses, Measurements) with all
    node ids 0 and pose ids the global g2o ids."""
    ii, jj, Rs, ts, kap, tau = [], [], [], [], [], []
    d = 0
    with open(filename) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "EDGE_SE2":
                d = 2
                i, j = int(tok[1]), int(tok[2])
                dx, dy, dth = (float(v) for v in tok[3:6])
                I11, I12, I13, I22, I23, I33 = (float(v) for v in tok[6:12])
                c, s = np.cos(dth), np.sin(dth)
                R = np.array([[c, -s], [s, c]])
                t = np.array([dx, dy])
                info = np.array([[I11, I12], [I12, I22]])
                tau_e = 2.0 / np.trace(np.linalg.inv(info))
                kap_e = I33
            elif tok[0] == "EDGE_SE3:QUAT":
                d = 3
                i, j = int(tok[1]), int(tok[2])
                v = [float(x) for x in tok[3:31]]
                dx, dy, dz, qx, qy, qz, qw = v[:7]
                I = v[7:]
                # upper-triangular 6x6, row-major
                I11, I12, I13 = I[0], I[1], I[2]
                I22, I23, I33 = I[6], I[7], I[11]
                I44, I45, I46 = I[15], I[16], I[17]
                I55, I56, I66 = I[18], I[19], I[20]
                R = quat_to_rot(qw, qx, qy, qz)
                t = np.array([dx, dy, dz])
                ti = np.array([[I11, I12, I13], [I12, I22, I23], [I13, I23, I33]])
                ri = np.array([[I44, I45, I46], [I45, I55, I56], [I46, I56, I66]])
                tau_e = 3.0 / np.trace(np.linalg.inv(ti))
                kap_e = 3.0 / (2.0 * np.trace(np.linalg.inv(ri)))
            elif tok[0] in ("VERTEX_SE2", "VERTEX_SE3:QUAT"):
                continue
            else:
                raise ValueError("unrecognized type: %s" % tok[0])
            ii.append(i); jj.append(j); Rs.append(R); ts.append(t)
            kap.append(kap_e); tau.append(tau_e)
    m = len(ii)
    num_poses = int(max(max(ii), max(jj))) + 1 if m else 0
    z = np.zeros(m, dtype=np.int64)
    return num_poses, Measurements(d, z, ii, z.copy(), jj, np.array(Rs),
                                   np.array(ts), kap, tau)


def partition_index(num_poses, num_nodes, gid):
    """The `index` lambda of read_g2o, DPGO_utils.cpp:147-158."""
    gid = np.asarray(gid, dtype=np.int64)
    q = num_poses // num_nodes
    inc_n = num_poses - num_nodes * q
    inc = inc_n * (q + 1)
    lo = gid < inc
    node = np.where(lo, gid // (q + 1), (gid - inc) // max(q, 1) + inc_n)
    pose = np.where(lo, gid % (q + 1), (gid - inc) % max(q, 1))
    return node, pose


def partition(num_poses, num_nodes, meas):
    """read_g2o, DPGO_utils.cpp:140-202.  Returns (per-node Measurements list
    with inter-node edges present in both endpoint lists, g_index) where
    g_index[a] is {local pose id: global id}."""
    ni, pi = partition_index(num_poses, num_nodes, meas.i_pose)
    nj, pj = partition_index(num_poses, num_nodes, meas.j_pose)
    part = Measurements(meas.d, ni, pi, nj, pj, meas.R, meas.t, meas.kappa,
                        meas.tau)
    per_node = []
    g_index = []
    for a in range(num_nodes):
        mask = (ni == a) | (nj == a)
        per_node.append(part.select(mask))
        g = {}
        for n_, p_, gid in ((ni, pi, meas.i_pose), (nj, pj, meas.j_pose)):
            sel = n_ == a
            for p, gg in zip(p_[sel], gid[sel]):
                g.setdefault(int(p), int(gg))
        g_index.append(dict(sorted(g.items())))
    return per_node, g_index, part
